// Internal interfaces between the C ABI layer, the Krylov drivers and the communication layer.
#pragma once
#include "common.cuh"

namespace b200 {

enum { B200_M_CG = 1, B200_M_BICGSTAB = 2, B200_M_BICGSTABL = 3, B200_M_GCR = 4, B200_M_IDRS = 5, B200_M_GMRES = 6, B200_M_CGS = 7, B200_M_TFQMR = 8, B200_M_BICGSTAB2 = 9, B200_M_JACOBI = 10, B200_M_RICHARDSON = 11, B200_M_SGS = 12 };

// b, x, P: device pointers; ipar/dpar: host HUTI arrays (fhutiter/src/huti_fdefs.h:101-155)
void solve_device(Handle &h, const double *d_b, double *d_x, int *ipar, double *dpar, int method, int pc, const double *d_P);

// in-place sum over ranks of `count` device doubles, on the solve stream (no-op for one rank)
void reduce_scalars(Handle &h, double *d, int count);
void comm_allreduce_sum(Handle &h, double *d, int count);

}  // namespace b200

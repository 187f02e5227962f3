// Host-side launchers of the kernel translation units (device pointers everywhere).
#pragma once
#include "common.cuh"

namespace b200 {

// ---- SpMV (spmv.cu) ------------------------------------------------------------------------
enum { EPI_NONE = 0, EPI_DOT1, EPI_DOT2, EPI_RESID, EPI_BMINUS };
struct SpmvArgs {
  const double *x = nullptr;   // operand (length n, plus ghosts when partitioned)
  double *y = nullptr;         // result
  const double *w = nullptr;   // EPI_DOT1/2: out[0] = sum y*w, out[1] = sum y*y
  const double *b = nullptr;   // EPI_RESID: out[0] = sum (Ax-b)^2 ; EPI_BMINUS: y = b - Ax, out[0] = sum y*y
  double *y2 = nullptr;        // EPI_BMINUS: optional second copy of y
  double *out = nullptr;       // device scalars receiving the reduction totals
  double *partials = nullptr; unsigned int *counter = nullptr;   // default: handle scratch
  const Ctrl *ctrl = nullptr;  // if set and ctrl->done, the kernel returns immediately
};
void spmv_launch(Handle &h, SpmvArgs a, int epi);   // owned x owned block only
void spmv_any(Handle &h, SpmvArgs a, int epi);      // + halo exchange and ghost block when partitioned (comm.cu)

// ---- BLAS-1 (blas1.cu) -----------------------------------------------------------------------
// out[k] = sum x_k[i]*y_k[i] for k < npairs (one pass, one reduction, deterministic)
void dot_batch(Handle &h, int n, int npairs, const double *const *xs, const double *const *ys, double *out);
void dot1(Handle &h, int n, const double *x, const double *y, double *out);
// y = a*x + b*y  (a, b host scalars; b == 0 never reads y)
void axpby(Handle &h, int n, double a, const double *x, double b, double *y);
// batched y_k = a_k*x_k + b_k*y_k in one launch (k < nops <= 8)
struct LinOp { const double *x; double *y; double a, b; };
void axpby_batch(Handle &h, int n, int nops, const LinOp *ops);
void copy_vec(Handle &h, int n, const double *x, double *y);
void div_scalar(Handle &h, int n, const double *x, double *y, double d);   // y = x / d (true division)
void fill_vec(Handle &h, long long n, double *x, double v);
// *flag = 1 if all x == 0 (device int)
void fill_if_all_zero(Handle &h, int n, double *x, double v);

}  // namespace b200

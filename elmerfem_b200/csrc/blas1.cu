// BLAS-1 building blocks used by the host-driven Krylov methods and the C ABI callbacks:
// batched dot products in one pass with one deterministic grid reduction (replaces ddot/dnrm2,
// mathlibs/src/blas/ddot.f, dnrm2.f as bound at fem/src/IterSolve.F90:910-912), batched
// y = a x + b y updates (the !$OMP PARALLEL DO vector loops of fem/src/IterativeMethods.F90).
// All streaming: one coalesced 8-byte access per thread and vector (256 B per warp request; measured at 0.9-1.0 of the HBM peak,
// so no 128-bit vectorisation), grid = multiple of 148.
#include "common.cuh"
#include "kernels.cuh"

namespace b200 {

struct DotArgs { const double *x[NRED]; const double *y[NRED]; int npairs; };

template <int NP>
__global__ void __launch_bounds__(256) k_dot_batch(int n, DotArgs a, double *partials, unsigned int *counter, double *out) {
  double acc[NP];
#pragma unroll
  for (int k = 0; k < NP; ++k) acc[k] = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
#pragma unroll
    for (int k = 0; k < NP; ++k) if (k < a.npairs) acc[k] += a.x[k][i] * a.y[k][i];
  }
  const int np = a.npairs;
  grid_reduce<NP>(acc, partials, counter, [out, np](double(&t)[NP]) {
#pragma unroll
    for (int k = 0; k < NP; ++k) if (k < np) out[k] = t[k];
  });
}

void dot_batch(Handle &h, int n, int npairs, const double *const *xs, const double *const *ys, double *out) {
  B200_REQUIRE(npairs >= 1 && npairs <= NRED, "dot_batch: too many pairs");
  DotArgs a; a.npairs = npairs;
  for (int k = 0; k < NRED; ++k) { a.x[k] = xs[k < npairs ? k : 0]; a.y[k] = ys[k < npairs ? k : 0]; }
  int blocks = std::max(1, std::min((n + 255) / 256, h.blas_blocks));
  cudaStream_t st = h.stream;
  if (npairs == 1) k_dot_batch<1><<<blocks, 256, 0, st>>>(n, a, h.red_partials.p, h.red_counters.p, out);
  else if (npairs == 2) k_dot_batch<2><<<blocks, 256, 0, st>>>(n, a, h.red_partials.p, h.red_counters.p, out);
  else if (npairs <= 4) k_dot_batch<4><<<blocks, 256, 0, st>>>(n, a, h.red_partials.p, h.red_counters.p, out);
  else if (npairs <= 8) k_dot_batch<8><<<blocks, 256, 0, st>>>(n, a, h.red_partials.p, h.red_counters.p, out);
  else if (npairs <= 16) k_dot_batch<16><<<blocks, 256, 0, st>>>(n, a, h.red_partials.p, h.red_counters.p, out);
  else k_dot_batch<NRED><<<blocks, 256, 0, st>>>(n, a, h.red_partials.p, h.red_counters.p, out);
  B200_CUDA(cudaGetLastError());
  h.st_launch++;
}
void dot1(Handle &h, int n, const double *x, const double *y, double *out) { dot_batch(h, n, 1, &x, &y, out); }

struct LinArgs { LinOp op[8]; int nops; };
__global__ void __launch_bounds__(256) k_axpby_batch(int n, LinArgs a) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < a.nops) {
        const LinOp &o = a.op[k];
        // separate roundings, evaluated as the reference's expressions a*x + b*y
        double ax = __dmul_rn(o.a, o.x[i]);
        double r = (o.b == 0.0) ? ax : ((o.b == 1.0) ? __dadd_rn(o.y[i], ax) : __dadd_rn(__dmul_rn(o.b, o.y[i]), ax));
        o.y[i] = r;
      }
    }
  }
}
void axpby_batch(Handle &h, int n, int nops, const LinOp *ops) {
  if (n == 0 || nops == 0) return;
  B200_REQUIRE(nops <= 8, "axpby_batch: too many ops");
  LinArgs a; a.nops = nops;
  for (int k = 0; k < 8; ++k) a.op[k] = ops[k < nops ? k : 0];
  int blocks = std::max(1, std::min((n + 255) / 256, h.blas_blocks));
  k_axpby_batch<<<blocks, 256, 0, h.stream>>>(n, a);
  B200_CUDA(cudaGetLastError());
  h.st_launch++;
}
void axpby(Handle &h, int n, double a, const double *x, double b, double *y) {
  LinOp o{x, y, a, b};
  axpby_batch(h, n, 1, &o);
}

// y = x / d with a true division (IDR(s) shadow space, IterativeMethods.F90:1659; GMRES basis vectors, huti_gmres.F90:189): x * (1/d)
// differs from x / d in the last bit
__global__ void __launch_bounds__(256) k_div_scalar(int n, const double *__restrict__ x, double *__restrict__ y, double d) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = __ddiv_rn(x[i], d);
}
void div_scalar(Handle &h, int n, const double *x, double *y, double d) {
  if (n == 0) return;
  int blocks = std::max(1, std::min((n + 255) / 256, h.blas_blocks));
  k_div_scalar<<<blocks, 256, 0, h.stream>>>(n, x, y, d);
  B200_CUDA(cudaGetLastError());
  h.st_launch++;
}

__global__ void k_copy(int n, const double *__restrict__ x, double *__restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = x[i];
}
void copy_vec(Handle &h, int n, const double *x, double *y) {
  if (n == 0 || x == y) return;
  B200_CUDA(cudaMemcpyAsync(y, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h.stream));
}

__global__ void k_fill_d(long long n, double *x, double v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = v;
}
void fill_vec(Handle &h, long long n, double *x, double v) {
  if (n == 0) return;
  int blocks = (int)std::max(1LL, std::min((n + 255) / 256, (long long)h.blas_blocks));
  k_fill_d<<<blocks, 256, 0, h.stream>>>(n, x, v);
  B200_CUDA(cudaGetLastError());
  h.st_launch++;
}

// IterSolve.F90:470-471: IF ( ALL(x == 0.0) ) x = 1.0d-8  (value passed in)
__global__ void k_any_nonzero(int n, const double *__restrict__ x, int *flag) {
  int nz = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) nz |= (x[i] != 0.0);
  if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}
__global__ void k_fill_if_flag0(int n, double *x, double v, const int *flag) {
  if (*flag) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = v;
}
void fill_if_all_zero(Handle &h, int n, double *x, double v) {
  if (n == 0) return;
  int *flag = &h.ctrl.p->flag;
  B200_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), h.stream));
  int blocks = std::max(1, std::min((n + 255) / 256, h.blas_blocks));
  k_any_nonzero<<<blocks, 256, 0, h.stream>>>(n, x, flag);
  k_fill_if_flag0<<<blocks, 256, 0, h.stream>>>(n, x, v, flag);
  B200_CUDA(cudaGetLastError());
  h.st_launch += 2;
}

}  // namespace b200

// Krylov drivers on device-resident vectors.
//
//   cg         <- huti_dcgsolv          fhutiter/src/huti_cg.F90:267-518        (keyword "cg")
//   bicgstab   <- huti_dbicgstabsolv    fhutiter/src/huti_bicgstab.F90:279-566  (keyword "bicgstab", default)
//   bicgstabl  <- RealBiCGStabl         fem/src/IterativeMethods.F90:694-1168
//   gcr        <- GCR                   fem/src/IterativeMethods.F90:1260-1458
//   idrs       <- RealIDRS              fem/src/IterativeMethods.F90:1579-1913
//
// CG and BiCGStab run without any host synchronisation inside an iteration: every scalar (rho, alpha,
// beta, omega, residual) lives in device memory, the dots and norms are fused into the SpMV epilogues
// and the vector-update kernels, the stopping test runs on the device and raises Ctrl::done, after
// which the kernels of iterations already queued return immediately.  The host polls the control
// block one iteration behind the GPU.  BiCGStab(l), GCR and IDR(s) keep their small dense algebra
// ((l+1)x(l+1) Gram system, s x s matrix M) on the host and fetch batched dot results once per step.
//
// The arithmetic of every vector update is written with the reference's association and separate
// multiply/add roundings; only the reductions (dot/norm) are summed in a different (tree) order.
#include "common.cuh"
#include "kernels.cuh"
#include "krylov.h"
#include <algorithm>
#include <cfloat>
#include <climits>

namespace b200 {

// device scalar slots; slots reduced together over ranks are adjacent (TS,TT) (RHONEXT,SN2)
enum { S_RHO = 0, S_OLDRHO, S_ALPHA, S_BETA, S_OMEGA, S_PQ, S_RTV, S_TS, S_TT, S_RHONEXT, S_SN2, S_RES2, S_BN2, S_TMP0 = 16 };

enum { HUTI_CONVERGENCE = 1, HUTI_MAXITER = 2, HUTI_DIVERGENCE = 3, HUTI_HALTED = 4,
       HUTI_CG_RHO = 20, HUTI_BICGSTAB_RHO = 35, HUTI_BICGSTAB_OMEGA = 37 };

#define IPAR(k) ipar[(k) - 1]
#define DPAR(k) dpar[(k) - 1]

// ---------------------------------------------------------------------------------------------
// shared device helpers
__device__ __forceinline__ double precond_elem(int pc, const double *dvals, int i, double v) {
  if (pc == 1) { double d = dvals[i]; return (fabs(d) > AEPS) ? __ddiv_rn(v, d) : v; }   // CRS_DiagPrecondition
  return v;
}

// residual of the selected stopping criterion (huti_cg.F90:408-464): stopc 0/2 unscaled, 1/3 over ||b||
__device__ __forceinline__ double stop_residual(const Ctrl *c, double sumsq) {
  double r = sqrt(sumsq);
  if (c->stopc == 1 || c->stopc == 3) r = r / c->bnorm;
  return r;
}

__global__ void k_init_ctrl(Ctrl *c, double *sc, double tol, double maxtol, int maxit, int minit, int stopc) {
  c->done = 0; c->info = 0; c->iters = 1; c->flag = 0; c->spin_timeout = 0; c->residual = 0.0;
  c->tol = tol; c->maxtol = maxtol; c->maxit = maxit; c->minit = minit; c->stopc = stopc; c->bnorm = 1.0;
  for (int i = 0; i < NSCAL; ++i) sc[i] = 0.0;
}
__global__ void k_set_bnorm(Ctrl *c, const double *sc) { c->bnorm = sqrt(sc[S_BN2]); }

// ---------------------------------------------------------------------------------------------
// CG
// P = Z (first) | P = Z + beta P, beta = rho/oldrho      huti_cg.F90:353-380
__global__ void __launch_bounds__(256) k_cg_p(int n, Ctrl *ctrl, double *sc, const double *__restrict__ z, double *__restrict__ p, int first) {
  if (ctrl->done) return;
  const double rho = sc[S_RHO];
  if (rho == 0.0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctrl->info = HUTI_CG_RHO; ctrl->done = 1; }
    return;
  }
  const double beta = first ? 0.0 : rho / sc[S_OLDRHO];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[i] = first ? z[i] : __dadd_rn(z[i], __dmul_rn(beta, p[i]));
}
// alpha = rho/(P.Q); X += alpha P; R -= alpha Q; [Z = M^-1 R for none/diagonal; rho' = R.Z]   384-402, 350-353
template <int PC>
__global__ void __launch_bounds__(256) k_cg_xr(int n, Ctrl *ctrl, double *sc, const double *__restrict__ p, const double *__restrict__ q,
                                                double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
                                                const double *__restrict__ dvals, double *partials, unsigned int *counter) {
  if (ctrl->done) return;
  const double alpha = sc[S_RHO] / sc[S_PQ];
  double acc[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    x[i] = __dadd_rn(x[i], __dmul_rn(alpha, p[i]));
    double ri = __dsub_rn(r[i], __dmul_rn(alpha, q[i]));
    r[i] = ri;
    acc[1] += ri * ri;
    if (PC < 2) {
      double zi = precond_elem(PC, dvals, i, ri);
      if (PC == 1) z[i] = zi;
      acc[0] += ri * zi;
    }
  }
  grid_reduce<2>(acc, partials, counter, [sc, alpha](double(&t)[2]) {
    sc[S_ALPHA] = alpha;
    if (PC < 2) sc[S_RHONEXT] = t[0];
    sc[S_SN2] = t[1];                    // ||R||^2, used by the pseudo-residual criteria
  });
}
// stopping test + bookkeeping of one CG iteration   huti_cg.F90:466-498
__global__ void k_cg_check(Ctrl *c, double *sc) {
  if (c->done) return;
  double sumsq = (c->stopc == 2 || c->stopc == 3) ? sc[S_SN2] : sc[S_RES2];
  double residual = stop_residual(c, sumsq);
  c->residual = residual;
  if (residual < c->tol) { c->info = HUTI_CONVERGENCE; c->done = 1; return; }
  if (residual != residual || residual > c->maxtol) { c->info = HUTI_DIVERGENCE; c->done = 1; return; }
  sc[S_OLDRHO] = sc[S_RHO];
  sc[S_RHO] = sc[S_RHONEXT];
  c->iters = c->iters + 1;
  if (c->iters > c->maxit) { c->info = HUTI_MAXITER; c->done = 1; }
}

// ---------------------------------------------------------------------------------------------
// BiCGStab
// beta = rho*alpha/(oldrho*omega); P = R + beta (P - omega V); [T1V = M^-1 P for diagonal]   382-401
template <int PC>
__global__ void __launch_bounds__(256) k_bicg_p(int n, Ctrl *ctrl, double *sc, const double *__restrict__ r, double *__restrict__ p,
                                                 const double *__restrict__ v, double *__restrict__ t1v, const double *__restrict__ dvals) {
  if (ctrl->done) return;
  const double rho = sc[S_RHO];
  if (rho == 0.0) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctrl->info = HUTI_BICGSTAB_RHO; ctrl->done = 1; }
    return;
  }
  const double omega = sc[S_OMEGA];
  const double beta = (rho * sc[S_ALPHA]) / (sc[S_OLDRHO] * omega);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double t = __dsub_rn(p[i], __dmul_rn(omega, v[i]));
    double pi = __dadd_rn(r[i], __dmul_rn(beta, t));
    p[i] = pi;
    if (PC == 1) t1v[i] = precond_elem(1, dvals, i, pi);
  }
}
// alpha = rho/(RTLD.V); S = R - alpha V; ||S||^2; [T2V = M^-1 S for diagonal]   404-415, 431-432
template <int PC>
__global__ void __launch_bounds__(256) k_bicg_s(int n, Ctrl *ctrl, double *sc, const double *__restrict__ r, const double *__restrict__ v,
                                                 double *__restrict__ s, double *__restrict__ t2v, const double *__restrict__ dvals,
                                                 double *partials, unsigned int *counter) {
  if (ctrl->done) return;
  const double alpha = sc[S_RHO] / sc[S_RTV];
  double acc[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double si = __dsub_rn(r[i], __dmul_rn(alpha, v[i]));
    s[i] = si;
    if (PC == 1) t2v[i] = precond_elem(1, dvals, i, si);
    acc[0] += si * si;
  }
  grid_reduce<1>(acc, partials, counter, [sc, alpha](double(&t)[1]) {
    sc[S_ALPHA] = alpha;
    sc[S_SN2] = t[0];
  });
}
// 415-429: if ||S|| < HUTI_EPSILON then X = X + alpha T1V and leave with HUTI_CONVERGENCE.  Runs after the
// (all-reduced) ||S||^2 is known; applied once, by the iteration `it` that detects it.
__global__ void __launch_bounds__(256) k_bicg_early(int n, Ctrl *ctrl, const double *sc, double *__restrict__ x,
                                                     const double *__restrict__ t1v, int it) {
  const int d = ctrl->done;
  if (d == 1 || (d == 2 && ctrl->iters != it)) return;
  if (!(sqrt(sc[S_SN2]) < HUTI_EPSILON)) return;
  const double alpha = sc[S_ALPHA];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    x[i] = __dadd_rn(x[i], __dmul_rn(alpha, t1v[i]));
  if (blockIdx.x == 0 && threadIdx.x == 0) { ctrl->info = HUTI_CONVERGENCE; ctrl->done = 2; }
}
// omega = (T.S)/(T.T); X += alpha T1V + omega T2V; R = S - omega T; rho' = RTLD.R ; ||R||^2   435-448, 382
__global__ void __launch_bounds__(256) k_bicg_xr(int n, Ctrl *ctrl, double *sc, double *__restrict__ x, const double *__restrict__ t1v,
                                                  const double *__restrict__ t2v, double *__restrict__ r, const double *__restrict__ s,
                                                  const double *__restrict__ t, const double *__restrict__ rtld,
                                                  double *partials, unsigned int *counter) {
  if (ctrl->done) return;
  const double alpha = sc[S_ALPHA];
  const double omega = sc[S_TS] / sc[S_TT];
  double acc[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    x[i] = __dadd_rn(__dadd_rn(x[i], __dmul_rn(alpha, t1v[i])), __dmul_rn(omega, t2v[i]));
    double ri = __dsub_rn(s[i], __dmul_rn(omega, t[i]));
    r[i] = ri;
    acc[0] += rtld[i] * ri;
    acc[1] += ri * ri;
  }
  grid_reduce<2>(acc, partials, counter, [sc, omega](double(&tt)[2]) {
    sc[S_OMEGA] = omega;
    sc[S_RHONEXT] = tt[0];
    sc[S_SN2] = tt[1];
  });
}
// huti_bicgstab.F90:512-547
__global__ void k_bicg_check(Ctrl *c, double *sc) {
  if (c->done) return;
  double sumsq = (c->stopc == 2 || c->stopc == 3) ? sc[S_SN2] : sc[S_RES2];
  double residual = stop_residual(c, sumsq);
  c->residual = residual;
  if (residual < c->tol) { c->info = HUTI_CONVERGENCE; c->done = 1; return; }
  if (sc[S_OMEGA] == 0.0) { c->info = HUTI_BICGSTAB_OMEGA; c->done = 1; return; }
  if (residual != residual || residual > c->maxtol) { c->info = HUTI_DIVERGENCE; c->done = 1; return; }
  sc[S_OLDRHO] = sc[S_RHO];
  sc[S_RHO] = sc[S_RHONEXT];
  c->iters = c->iters + 1;
  if (c->iters > c->maxit) { c->info = HUTI_MAXITER; c->done = 1; }
}
__global__ void k_bicg_init_scalars(double *sc) {
  sc[S_OLDRHO] = 1.0; sc[S_OMEGA] = 1.0; sc[S_ALPHA] = 0.0;      // huti_bicgstab.F90:373
  sc[S_RHO] = sc[S_TMP0];                                        // RTLD.R = R.R of the initial residual
}
__global__ void k_cg_init_scalars(double *sc) { sc[S_RHO] = sc[S_RHONEXT]; }

// ---------------------------------------------------------------------------------------------
struct Solver {
  Handle &h; int n; int pc; cudaStream_t st; Ctrl *ctrl; double *sc; int blocks;
  std::vector<double *> vec;
  Solver(Handle &h_, int pc_, int nvec) : h(h_), n(h_.n), pc(pc_), st(h_.stream), ctrl(h_.ctrl.p), sc(h_.scal.p) {
    blocks = std::max(1, std::min((n + 255) / 256, h.blas_blocks));
    size_t stride = (vec_len(h) + 31) / 32 * 32 + 32;       // room for the ghost tail of SpMV operands
    if (h.work.empty()) h.work.resize(1);
    h.work[0].ensure(stride * nvec);
    for (int k = 0; k < nvec; ++k) vec.push_back(h.work[0].p + stride * k);
    B200_CUDA(cudaMemsetAsync(h.work[0].p, 0, stride * nvec * sizeof(double), st));   // IterSolve.F90:455-467 work = 0
  }
  // u = M^-1 v through the selected right preconditioner (IterSolve.F90:816-836); returns the vector holding u
  double *precond(double *u, double *v) {
    if (pc == 2) { lu_apply(h, u, v); return u; }
    if (pc == 1) { diag_apply(h, u, v); h.st_pcond++; return u; }
    h.st_pcond++;
    return v;                          // pcond_dummy: u = v, elided by aliasing
  }
  void matvec(const double *x, double *y) { matvec_full(h, x, y); }
  // batched host-visible dots
  void dots(int np, const double *const *xs, const double *const *ys, double *host_out) {
    dot_batch(h, n, np, xs, ys, sc + S_TMP0);
    reduce_scalars(h, sc + S_TMP0, np);
    B200_CUDA(cudaMemcpyAsync(h.h_pinned, sc + S_TMP0, np * sizeof(double), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < np; ++k) host_out[k] = h.h_pinned[k];
    h.st_d2h += np * sizeof(double);
  }
  double dot(const double *x, const double *y) { double r; dots(1, &x, &y, &r); return r; }
  double norm(const double *x) { return sqrt(dot(x, x)); }
  void lin(const double *x, double a, double *y, double b) { axpby(h, n, a, x, b, y); }    // y = a x + b y
};

void reduce_scalars(Handle &h, double *d, int count) {
  if (h.nranks > 1) comm_allreduce_sum(h, d, count);
}

static void read_ctrl(Handle &h) {
  B200_CUDA(cudaMemcpyAsync(h.h_ctrl, h.ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, h.stream));
  B200_CUDA(cudaStreamSynchronize(h.stream));
}

// Polls the control block one iteration behind the launch front: the copy of iteration `it` is
// queued, then the host waits for the copy of iteration it-1.
struct Poller {
  Handle &h; cudaEvent_t ev[2]; Ctrl *slot[2]; bool pending[2] = {false, false};
  Poller(Handle &h_) : h(h_) {
    ev[0] = h.ev1; ev[1] = h.ev2; slot[0] = h.h_ctrl; slot[1] = h.h_ctrl + 1;
  }
  // returns true when an already-finished iteration reported done
  bool step(int it) {
    int cur = it & 1, prev = cur ^ 1;
    B200_CUDA(cudaMemcpyAsync(slot[cur], h.ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, h.stream));
    B200_CUDA(cudaEventRecord(ev[cur], h.stream));
    pending[cur] = true;
    if (pending[prev]) {
      B200_CUDA(cudaEventSynchronize(ev[prev]));
      pending[prev] = false;
      if (slot[prev]->done || slot[prev]->spin_timeout) return true;
    }
    return false;
  }
};

// ---------------------------------------------------------------------------------------------
static void run_cg(Handle &h, const double *b, double *x, int pc, int maxit, int stopc) {
  const bool true_resid = (stopc == 0 || stopc == 1);
  Solver S(h, pc, 4);
  double *Z = S.vec[0], *P = S.vec[1], *Q = S.vec[2], *R = S.vec[3];
  const int n = S.n; cudaStream_t st = S.st; Ctrl *ctrl = S.ctrl; double *sc = S.sc;
  double *zz = (pc == 0) ? R : Z;                       // Z aliases R without a preconditioner
  // rhsnorm = ||B|| (310-313); R = B - A X (331-341)
  dot1(h, n, b, b, sc + S_BN2); reduce_scalars(h, sc + S_BN2, 1);
  k_set_bnorm<<<1, 1, 0, st>>>(ctrl, sc);
  { SpmvArgs a; a.x = x; a.y = R; a.b = b; a.out = sc + S_TMP0; spmv_any(h, a, EPI_BMINUS); }
  // Z = M^-1 R ; rho = R.Z  (350-353) for the first pass
  S.precond(Z, R);
  dot1(h, n, R, zz, sc + S_RHONEXT); reduce_scalars(h, sc + S_RHONEXT, 1);
  k_cg_init_scalars<<<1, 1, 0, st>>>(sc);
  h.st_launch += 2;
  Poller poll(h);
  for (int it = 1; it <= maxit + 1; ++it) {
    k_cg_p<<<S.blocks, 256, 0, st>>>(n, ctrl, sc, zz, P, it == 1 ? 1 : 0);
    { SpmvArgs a; a.x = P; a.y = Q; a.w = P; a.out = sc + S_PQ; a.ctrl = ctrl; spmv_any(h, a, EPI_DOT1); reduce_scalars(h, sc + S_PQ, 1); }
    if (pc == 0) k_cg_xr<0><<<S.blocks, 256, 0, st>>>(n, ctrl, sc, P, Q, x, R, Z, h.d_dvals.p, h.red_partials.p, h.red_counters.p);
    else if (pc == 1) k_cg_xr<1><<<S.blocks, 256, 0, st>>>(n, ctrl, sc, P, Q, x, R, Z, h.d_dvals.p, h.red_partials.p, h.red_counters.p);
    else k_cg_xr<2><<<S.blocks, 256, 0, st>>>(n, ctrl, sc, P, Q, x, R, Z, h.d_dvals.p, h.red_partials.p, h.red_counters.p);
    reduce_scalars(h, sc + S_RHONEXT, 2);
    if (pc == 2) {                                   // next pass's Z = (LU)^-1 R and rho = R.Z
      lu_apply(h, Z, R);
      dot1(h, n, R, Z, sc + S_RHONEXT); reduce_scalars(h, sc + S_RHONEXT, 1);
    } else h.st_pcond++;
    // true residual ||A X - B|| (421-432); skipped by the pseudo-residual criteria
    if (true_resid) { SpmvArgs a; a.x = x; a.b = b; a.out = sc + S_RES2; a.ctrl = ctrl; spmv_any(h, a, EPI_RESID); reduce_scalars(h, sc + S_RES2, 1); }
    k_cg_check<<<1, 1, 0, st>>>(ctrl, sc);
    h.st_launch += 3;
    B200_CUDA(cudaGetLastError());
    if (poll.step(it)) break;
  }
  read_ctrl(h);
}

static void run_bicgstab(Handle &h, const double *b, double *x, int pc, int maxit, int stopc) {
  const bool true_resid = (stopc == 0 || stopc == 1);
  Solver S(h, pc, 8);
  double *RTLD = S.vec[0], *P = S.vec[1], *T1V = S.vec[2], *V = S.vec[3], *Sv = S.vec[4], *T2V = S.vec[5], *T = S.vec[6], *R = S.vec[7];
  const int n = S.n; cudaStream_t st = S.st; Ctrl *ctrl = S.ctrl; double *sc = S.sc;
  double *t1 = (pc == 0) ? P : T1V, *t2 = (pc == 0) ? Sv : T2V;     // dummy preconditioner elided by aliasing
  dot1(h, n, b, b, sc + S_BN2); reduce_scalars(h, sc + S_BN2, 1);
  k_set_bnorm<<<1, 1, 0, st>>>(ctrl, sc);
  // R = B - A X, RTLD = R (352-359); P = V = 0 from the zeroed work array; rho = RTLD.R
  { SpmvArgs a; a.x = x; a.y = R; a.y2 = RTLD; a.b = b; a.out = sc + S_TMP0; spmv_any(h, a, EPI_BMINUS); reduce_scalars(h, sc + S_TMP0, 1); }
  k_bicg_init_scalars<<<1, 1, 0, st>>>(sc);
  h.st_launch += 2;
  Poller poll(h);
  for (int it = 1; it <= maxit + 1; ++it) {
    if (pc == 1) k_bicg_p<1><<<S.blocks, 256, 0, st>>>(n, ctrl, sc, R, P, V, T1V, h.d_dvals.p);
    else k_bicg_p<0><<<S.blocks, 256, 0, st>>>(n, ctrl, sc, R, P, V, T1V, h.d_dvals.p);
    if (pc == 2) lu_apply(h, T1V, P); else h.st_pcond++;
    { SpmvArgs a; a.x = t1; a.y = V; a.w = RTLD; a.out = sc + S_RTV; a.ctrl = ctrl; spmv_any(h, a, EPI_DOT1); reduce_scalars(h, sc + S_RTV, 1); }
    if (pc == 1) k_bicg_s<1><<<S.blocks, 256, 0, st>>>(n, ctrl, sc, R, V, Sv, T2V, h.d_dvals.p, h.red_partials.p, h.red_counters.p);
    else k_bicg_s<0><<<S.blocks, 256, 0, st>>>(n, ctrl, sc, R, V, Sv, T2V, h.d_dvals.p, h.red_partials.p, h.red_counters.p);
    reduce_scalars(h, sc + S_SN2, 1);
    k_bicg_early<<<S.blocks, 256, 0, st>>>(n, ctrl, sc, x, t1, it);
    if (pc == 2) lu_apply(h, T2V, Sv); else h.st_pcond++;
    { SpmvArgs a; a.x = t2; a.y = T; a.w = Sv; a.out = sc + S_TS; a.ctrl = ctrl; spmv_any(h, a, EPI_DOT2); reduce_scalars(h, sc + S_TS, 2); }
    k_bicg_xr<<<S.blocks, 256, 0, st>>>(n, ctrl, sc, x, t1, t2, R, Sv, T, RTLD, h.red_partials.p, h.red_counters.p);
    reduce_scalars(h, sc + S_RHONEXT, 2);
    if (true_resid) { SpmvArgs a; a.x = x; a.b = b; a.out = sc + S_RES2; a.ctrl = ctrl; spmv_any(h, a, EPI_RESID); reduce_scalars(h, sc + S_RES2, 1); }
    k_bicg_check<<<1, 1, 0, st>>>(ctrl, sc);
    h.st_launch += 5;
    B200_CUDA(cudaGetLastError());
    if (poll.step(it)) break;
  }
  read_ctrl(h);
  if (h.h_ctrl->done == 2) h.h_ctrl->done = 1;
}

// ---------------------------------------------------------------------------------------------
// small dense helpers for BiCGStab(l): partial-pivot LU (dgetrf/dgetrs) and dsymv('u'), column major
struct SmallLU {
  int n = 0; std::vector<double> a; std::vector<int> piv;
  void factor(int n_, const double *A, int lda) {
    n = n_; a.assign((size_t)n * n, 0.0); piv.assign(n, 0);
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) a[i + (size_t)j * n] = A[i + (size_t)j * lda];
    for (int j = 0; j < n; ++j) {
      int p = j; double mx = fabs(a[j + (size_t)j * n]);
      for (int i = j + 1; i < n; ++i) if (fabs(a[i + (size_t)j * n]) > mx) { mx = fabs(a[i + (size_t)j * n]); p = i; }
      piv[j] = p;
      if (a[p + (size_t)j * n] != 0.0) {
        if (p != j) for (int k = 0; k < n; ++k) std::swap(a[j + (size_t)k * n], a[p + (size_t)k * n]);
        double r = 1.0 / a[j + (size_t)j * n];
        for (int i = j + 1; i < n; ++i) a[i + (size_t)j * n] *= r;
      }
      for (int k = j + 1; k < n; ++k) for (int i = j + 1; i < n; ++i) a[i + (size_t)k * n] -= a[i + (size_t)j * n] * a[j + (size_t)k * n];
    }
  }
  void solve(double *b) const {
    for (int j = 0; j < n; ++j) if (piv[j] != j) std::swap(b[j], b[piv[j]]);
    for (int j = 0; j < n; ++j) for (int i = j + 1; i < n; ++i) b[i] -= b[j] * a[i + (size_t)j * n];
    for (int j = n - 1; j >= 0; --j) { b[j] /= a[j + (size_t)j * n]; for (int i = 0; i < j; ++i) b[i] -= b[j] * a[i + (size_t)j * n]; }
  }
};
static void symv_u(int n, const double *A, int lda, const double *x, double *y) {
  for (int i = 0; i < n; ++i) y[i] = 0.0;
  for (int j = 0; j < n; ++j) {
    double t1 = x[j], t2 = 0.0;
    for (int i = 0; i < j; ++i) { y[i] += t1 * A[i + (size_t)j * lda]; t2 += A[i + (size_t)j * lda] * x[i]; }
    y[j] += t1 * A[j + (size_t)j * lda] + t2;
  }
}
static double sdot(int n, const double *x, const double *y) { double s = 0; for (int i = 0; i < n; ++i) s += x[i] * y[i]; return s; }

struct HostResult { int info = 0, iters = 0; double residual = 0; };
// `Linear System Robust` (IterSolve.F90:482-496): keep the best iterate seen and stop on it when the iteration has
// wandered off; in the reference only BiCGStab(l) and IDR(s) look at it (IterativeMethods.F90:649-659, 1535-1544)
struct RobustPar { bool on = false; double Tol = 0, Step = 0, MaxTol = 0; int MaxBadIter = 0, Start = 1; };

// IterativeMethods.F90:694-1168
static HostResult run_bicgstabl(Handle &h, const double *b, double *x, int pc, int MaxRounds, double Tol, double MaxTol, int l,
                                const RobustPar &Rb) {
  HostResult res;
  B200_REQUIRE(l >= 2, "BiCGStab(l): polynomial degree < 2");
  const int nw = 3 + 2 * (l + 1);
  Solver S(h, pc, nw + 1 + (Rb.on ? 1 : 0));
  const int n = S.n;
  auto work = [&](int c) { return S.vec[c - 1]; };
  double *t = S.vec[nw];
  double *Bestx = Rb.on ? S.vec[nw + 1] : nullptr;              // 649-659
  double BestNorm = sqrt(DBL_MAX);
  int BadIterCount = 0;
  const int rr = 1, r = rr + 1, u = r + (l + 1), xp = u + (l + 1), bp = xp + 1;
  const int ldr = l + 1;
  std::vector<double> rw((size_t)ldr * nw, 0.0);
  auto rwork = [&](int i, int j) -> double & { return rw[(i - 1) + (size_t)(j - 1) * ldr]; };
  const int z = 1, zz = z + (l + 1), y0 = zz + (l + 1), yl = y0 + 1, y = yl + 1;
  std::vector<double> tmpmtr((size_t)(l - 1) * (l - 1)), tmpvec(l - 1);
  SmallLU lu;
  bool Converged = false, Diverged = false, Halted = false;
  // 719: IF ( ALL(x == 0) ) x = b -- unreachable after IterSolver's x = 1e-8 rule unless b == 0; kept
  {
    double nx2 = S.dot(x, x);
    if (nx2 == 0.0) copy_vec(h, n, b, x);
  }
  S.matvec(x, work(r));
  S.lin(b, 1.0, work(r), -1.0);                               // r = b - r
  double bnrm, rnrm0;
  { const double *xs[2] = {b, work(r)}, *ys[2] = {b, work(r)}; double o[2]; S.dots(2, xs, ys, o); bnrm = sqrt(o[0]); rnrm0 = sqrt(o[1]); }
  double errorind = rnrm0 / bnrm;
  if (bnrm != bnrm || rnrm0 != rnrm0 || errorind != errorind) { res.info = HUTI_DIVERGENCE; res.residual = errorind; return res; }
  Converged = errorind < Tol; Diverged = errorind > MaxTol;
  int Round = 0;
  if (Converged || Diverged) { res.info = Converged ? HUTI_CONVERGENCE : HUTI_DIVERGENCE; res.residual = errorind; return res; }
  copy_vec(h, n, work(r), work(rr)); copy_vec(h, n, work(r), work(bp));
  copy_vec(h, n, x, work(xp));
  fill_vec(h, n, x, 0.0);
  double rnrm = rnrm0, mxnrmx = rnrm0, mxnrmr = rnrm0;
  double alpha = 0.0, omega = 1.0, sigma = 1.0, rho0 = 1.0, rho1, beta;
  bool EarlyExit = false;
  for (Round = 1; Round <= MaxRounds; ++Round) {
    rho0 = -omega * rho0;
    for (int k = 1; k <= l; ++k) {
      rho1 = S.dot(work(rr), work(r + k - 1));
      if (rho0 == 0.0) { Halted = true; goto L100; }
      if (rho1 != rho1) { Diverged = true; goto L100; }
      beta = alpha * (rho1 / rho0);
      rho0 = rho1;
      for (int j0 = 0; j0 <= k - 1; j0 += 8) {                 // u_j = r_j - beta u_j, j < k   (836-842)
        LinOp ops[8]; int m = 0;
        for (int j = j0; j <= k - 1 && m < 8; ++j) ops[m++] = LinOp{work(r + j), work(u + j), 1.0, -beta};
        axpby_batch(h, n, m, ops);
      }
      { double *tt = S.precond(t, work(u + k - 1)); S.matvec(tt, work(u + k)); }
      sigma = S.dot(work(rr), work(u + k));
      if (sigma == 0.0) { Halted = true; goto L100; }
      if (sigma != sigma) { Diverged = true; goto L100; }
      alpha = rho1 / sigma;
      {                                                        // x += alpha u_0 ; r_j -= alpha u_{j+1}  (865-875)
        LinOp ops[8]; int m = 0;
        ops[m++] = LinOp{work(u), x, alpha, 1.0};
        for (int j = 0; j <= k - 1; ++j) {
          ops[m++] = LinOp{work(u + j + 1), work(r + j), -alpha, 1.0};
          if (m == 8) { axpby_batch(h, n, m, ops); m = 0; }
        }
        if (m) axpby_batch(h, n, m, ops);
      }
      { double *tt = S.precond(t, work(r + k - 1)); S.matvec(tt, work(r + k)); }
      rnrm = S.norm(work(r));
      if (rnrm != rnrm) { Diverged = true; goto L100; }
      mxnrmx = std::max(mxnrmx, rnrm); mxnrmr = std::max(mxnrmr, rnrm);
      errorind = rnrm / bnrm;
      Converged = errorind < Tol; Diverged = errorind != errorind;
      if (Converged || Diverged) { EarlyExit = true; break; }
    }
    if (EarlyExit) break;
    {                                                          // Gram matrix, one batched pass (917-922)
      std::vector<const double *> xs, ys; std::vector<double> o((l + 1) * (l + 2) / 2);
      for (int i = 1; i <= l + 1; ++i) for (int j = 1; j <= i; ++j) { xs.push_back(work(r + i - 1)); ys.push_back(work(r + j - 1)); }
      int done = 0, tot = (int)xs.size();
      while (done < tot) { int m = std::min(NRED, tot - done); S.dots(m, xs.data() + done, ys.data() + done, o.data() + done); done += m; }
      int q = 0;
      for (int i = 1; i <= l + 1; ++i) for (int j = 1; j <= i; ++j) rwork(i, j) = o[q++];
    }
    for (int j = 2; j <= l + 1; ++j) for (int i = 1; i <= j - 1; ++i) rwork(i, j) = rwork(j, i);
    for (int j = 0; j <= l - 1; ++j) for (int i = 1; i <= l + 1; ++i) rwork(i, zz + j) = rwork(i, z + j);
    for (int j = 1; j <= l - 1; ++j) for (int i = 1; i <= l - 1; ++i) tmpmtr[(i - 1) + (size_t)(j - 1) * (l - 1)] = rwork(i + 1, zz + j);
    lu.factor(l - 1, tmpmtr.data(), l - 1);
    rwork(1, y0) = -1.0;
    for (int i = 2; i <= l; ++i) rwork(i, y0) = rwork(i, z);
    for (int i = 1; i <= l - 1; ++i) tmpvec[i - 1] = rwork(i + 1, y0);
    lu.solve(tmpvec.data());
    for (int i = 1; i <= l - 1; ++i) rwork(i + 1, y0) = tmpvec[i - 1];
    rwork(l + 1, y0) = 0.0;
    rwork(1, yl) = 0.0;
    for (int i = 1; i <= l - 1; ++i) { rwork(i + 1, yl) = rwork(i + 1, z + l); tmpvec[i - 1] = rwork(i + 1, yl); }
    lu.solve(tmpvec.data());
    for (int i = 1; i <= l - 1; ++i) rwork(i + 1, yl) = tmpvec[i - 1];
    rwork(l + 1, yl) = -1.0;
    {
      double kappa0, kappal, varrho, hatgamma;
      symv_u(l + 1, &rwork(1, z), ldr, &rwork(1, y0), &rwork(1, y));
      kappa0 = sdot(l + 1, &rwork(1, y0), &rwork(1, y));
      if (kappa0 <= 0.0) { Halted = true; goto L100; }
      kappa0 = sqrt(kappa0);
      symv_u(l + 1, &rwork(1, z), ldr, &rwork(1, yl), &rwork(1, y));
      kappal = sdot(l + 1, &rwork(1, yl), &rwork(1, y));
      if (kappal <= 0.0) { Halted = true; goto L100; }
      kappal = sqrt(kappal);
      symv_u(l + 1, &rwork(1, z), ldr, &rwork(1, y0), &rwork(1, y));
      varrho = sdot(l + 1, &rwork(1, yl), &rwork(1, y)) / (kappa0 * kappal);
      hatgamma = varrho / fabs(varrho) * std::max(fabs(varrho), 7e-1) * kappa0 / kappal;
      for (int i = 1; i <= l + 1; ++i) rwork(i, y0) = rwork(i, y0) - hatgamma * rwork(i, yl);
    }
    omega = rwork(l + 1, y0);
    for (int j = 1; j <= l; ++j) {                             // 1016-1032, one launch per j (three updates)
      double g = rwork(j + 1, y0);
      LinOp ops[3] = {LinOp{work(u + j), work(u), -g, 1.0}, LinOp{work(r + j - 1), x, g, 1.0}, LinOp{work(r + j), work(r), -g, 1.0}};
      axpby_batch(h, n, 3, ops);   // per element the three updates run in the reference's order
    }
    symv_u(l + 1, &rwork(1, z), ldr, &rwork(1, y0), &rwork(1, y));
    rnrm = sdot(l + 1, &rwork(1, y0), &rwork(1, y));
    if (rnrm < 0.0) { Halted = true; goto L100; }
    rnrm = sqrt(rnrm);
    {                                                          // reliable update (1050-1101)
      mxnrmx = std::max(mxnrmx, rnrm); mxnrmr = std::max(mxnrmr, rnrm);
      bool xpdt = (rnrm < 1.0e-2 * rnrm0 && rnrm0 < mxnrmx);
      bool rcmp = ((rnrm < 1.0e-2 * mxnrmr && rnrm0 < mxnrmr) || xpdt);
      if (rcmp) {
        double *tt = S.precond(t, x);
        S.matvec(tt, work(r));
        mxnrmr = rnrm;
        S.lin(work(bp), 1.0, work(r), -1.0);                   // r = bp - r
        if (xpdt) {
          S.lin(tt, 1.0, work(xp), 1.0);                       // xp += t
          fill_vec(h, n, x, 0.0);
          copy_vec(h, n, work(r), work(bp));
          mxnrmx = rnrm;
        }
      }
      // 1080-1101 only produce a vector t that is never read again (one dead preconditioner solve per
      // round when rcmp is false): not executed, numbers unchanged.
    }
    errorind = rnrm / bnrm;
    if (Rb.on && Round >= Rb.Start) {                          // 1110-1126
      if (errorind < Rb.Step * BestNorm) { BestNorm = errorind; copy_vec(h, n, x, Bestx); BadIterCount = 0; }
      else BadIterCount = BadIterCount + 1;
      if (BestNorm < Rb.Tol && (errorind > Rb.MaxTol || BadIterCount > Rb.MaxBadIter)) break;
    }
    Converged = errorind < Tol;
    Diverged = (errorind > MaxTol) || (errorind != errorind);
    if (Converged || Diverged) break;
  }
L100:
  if (Rb.on) {                                                 // 1133-1139
    if (BestNorm < Rb.Tol) Converged = true;
    if (BestNorm < errorind) copy_vec(h, n, Bestx, x);
  }
  res.iters = std::min(MaxRounds, Round);
  res.residual = errorind;
  // 1156-1166: x = M^-1 x + xp
  copy_vec(h, n, x, t);
  if (pc != 0) { S.precond(x, t); } else h.st_pcond++;
  S.lin(work(xp), 1.0, x, 1.0);
  if (Converged) res.info = HUTI_CONVERGENCE;
  else if (Diverged) res.info = HUTI_DIVERGENCE;
  else if (Halted) res.info = HUTI_HALTED;
  else res.info = HUTI_MAXITER;
  return res;
}

// ---------------------------------------------------------------------------------------------
// BiCGStab(l), device resident (the default; `Linear System Robust` keeps the host-driven driver above).
// Every scalar of RealBiCGStabl lives in device memory; the (l+1) x (l+1) Gram algebra of the convex-polynomial part (940-1037: dgetrf /
// dgetrs / dsymv / ddot on the small work array) runs in ONE thread, restated operation by operation; the dots of the BiCG part ride in
// the epilogues of the SpMVs that produce their operand (sigma = rr.u_k with u_k = A M^-1 u_(k-1); the next rho1 = rr.r_k with
// r_k = A M^-1 r_(k-1)) and ||r_0||^2 in the kernel that updates r_0, so a round is a fixed sequence of launches with no host
// synchronisation: the stopping tests raise Ctrl::done, after which already-queued kernels return at once; the host polls the pinned
// control block one round behind.  The reliable-update branch (1050-1079: recompute r = b' - A M^-1 x, flying restart) is queued every
// round and skipped on the device (Ctrl::done = 2 for the length of the section) when `rcmp` is false.
// Reductions over ranks per round: 2 per k (sigma; ||r_0||^2 stacked with the next rho1) + the Gram matrix = 2 l + 1, each a true
// dependency of the next vector update.
enum { BL_RHO0 = 0, BL_RHO1S, BL_ALPHA, BL_BETA, BL_OMEGA, BL_SIGMA, BL_RNRM2, BL_RHO1, BL_RNRM, BL_RNRM0, BL_BNRM, BL_MXX, BL_MXR, BL_XPDT, BL_RCMP, BL_ROUND,
       BL_GRAM = 16, BL_GAM = 40 };
constexpr int BL_MAXL = 5;       // (l+1)(l+2)/2 <= NRED Gram dots in one pass

__global__ void k_bl_init(double *sc, double bnrm, double rnrm0) {
  sc[BL_RHO0] = 1.0; sc[BL_ALPHA] = 0.0; sc[BL_OMEGA] = 1.0; sc[BL_SIGMA] = 1.0;
  sc[BL_RNRM] = rnrm0; sc[BL_RNRM0] = rnrm0; sc[BL_BNRM] = bnrm; sc[BL_MXX] = rnrm0; sc[BL_MXR] = rnrm0; sc[BL_XPDT] = 0.0; sc[BL_RCMP] = 0.0; sc[BL_ROUND] = 1.0;
}
// 820-834: [rho0 = -omega rho0 at the top of a round]; beta = alpha (rho1 / rho0); rho0 = rho1
__global__ void k_bl_rho(Ctrl *c, double *sc, int first) {
  if (c->done) return;
  double rho0 = sc[BL_RHO0];
  if (first) rho0 = -sc[BL_OMEGA] * rho0;
  const double rho1 = sc[BL_RHO1];
  c->iters = (int)sc[BL_ROUND];
  if (rho0 == 0.0) { c->info = HUTI_HALTED; c->done = 1; return; }
  if (rho1 != rho1) { c->info = HUTI_DIVERGENCE; c->done = 1; return; }
  sc[BL_BETA] = sc[BL_ALPHA] * (rho1 / rho0);
  sc[BL_RHO0] = rho1; sc[BL_RHO1S] = rho1;
}
struct BlVecs { double *r[BL_MAXL + 1]; double *u[BL_MAXL + 1]; };
// u_j = r_j - beta u_j, j < k   (836-842)
__global__ void __launch_bounds__(256) k_bl_uupd(int n, const Ctrl *c, const double *sc, BlVecs v, int k) {
  if (c->done) return;
  const double beta = sc[BL_BETA];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    for (int j = 0; j < k; ++j) v.u[j][i] = __dsub_rn(v.r[j][i], __dmul_rn(beta, v.u[j][i]));
}
// 852-862: alpha = rho1 / sigma
__global__ void k_bl_alpha(Ctrl *c, double *sc) {
  if (c->done) return;
  const double sigma = sc[BL_SIGMA];
  if (sigma == 0.0) { c->info = HUTI_HALTED; c->done = 1; return; }
  if (sigma != sigma) { c->info = HUTI_DIVERGENCE; c->done = 1; return; }
  sc[BL_ALPHA] = sc[BL_RHO1S] / sigma;
}
// x += alpha u_0 ; r_j -= alpha u_(j+1), j < k ; ||r_0||^2   (865-875, 886)
__global__ void __launch_bounds__(256) k_bl_xr(int n, const Ctrl *c, double *sc, double *__restrict__ x, BlVecs v, int k, double *partials, unsigned int *counter) {
  if (c->done) return;
  const double alpha = sc[BL_ALPHA];
  double acc[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    x[i] = __dadd_rn(x[i], __dmul_rn(alpha, v.u[0][i]));
    for (int j = 0; j < k; ++j) {
      const double rj = __dsub_rn(v.r[j][i], __dmul_rn(alpha, v.u[j + 1][i]));
      v.r[j][i] = rj;
      if (j == 0) acc[0] += rj * rj;
    }
  }
  grid_reduce<1>(acc, partials, counter, [sc](double(&t)[1]) { sc[BL_RNRM2] = t[0]; });
}
// 886-909: rnrm = ||r_0||, running maxima, convergence inside the BiCG part (EarlyExit)
__global__ void k_bl_check(Ctrl *c, double *sc, double Tol) {
  if (c->done) return;
  const double rnrm = sqrt(sc[BL_RNRM2]);
  sc[BL_RNRM] = rnrm;
  if (rnrm != rnrm) { c->info = HUTI_DIVERGENCE; c->done = 1; return; }
  sc[BL_MXX] = fmax(sc[BL_MXX], rnrm); sc[BL_MXR] = fmax(sc[BL_MXR], rnrm);
  const double errorind = rnrm / sc[BL_BNRM];
  c->residual = errorind;
  if (errorind < Tol) { c->info = HUTI_CONVERGENCE; c->done = 1; return; }
  if (errorind != errorind) { c->info = HUTI_DIVERGENCE; c->done = 1; return; }
}
// 917-1011, 1036-1068: the convex polynomial part on the Gram matrix and the reliable-update decisions, one thread
template <int L>
__global__ void k_bl_poly(Ctrl *c, double *sc) {
  if (c->done) return;
  constexpr int ldr = L + 1, nw = 3 + 2 * (L + 1);
  double rw[ldr * nw];
  for (int q = 0; q < ldr * nw; ++q) rw[q] = 0.0;
#define RW(i, j) rw[((i) - 1) + ((j) - 1) * ldr]
  constexpr int z = 1, zz = z + (L + 1), y0 = zz + (L + 1), yl = y0 + 1, y = yl + 1;
  { int q = 0; for (int i = 1; i <= L + 1; ++i) for (int j = 1; j <= i; ++j) RW(i, j) = sc[BL_GRAM + q++]; }
  for (int j = 2; j <= L + 1; ++j) for (int i = 1; i <= j - 1; ++i) RW(i, j) = RW(j, i);
  for (int j = 0; j <= L - 1; ++j) for (int i = 1; i <= L + 1; ++i) RW(i, zz + j) = RW(i, z + j);
  // dgetrf on rwork(2:l, zz+1:zz+l-1): partial pivoting, unit lower
  constexpr int m = L - 1;
  double a[m * m]; int piv[m]; double tv[m];
  for (int j = 1; j <= m; ++j) for (int i = 1; i <= m; ++i) a[(i - 1) + (j - 1) * m] = RW(i + 1, zz + j);
  for (int j = 0; j < m; ++j) {
    int p = j; double mx = fabs(a[j + j * m]);
    for (int i = j + 1; i < m; ++i) if (fabs(a[i + j * m]) > mx) { mx = fabs(a[i + j * m]); p = i; }
    piv[j] = p;
    if (a[p + j * m] != 0.0) {
      if (p != j) for (int k = 0; k < m; ++k) { const double t = a[j + k * m]; a[j + k * m] = a[p + k * m]; a[p + k * m] = t; }
      const double r = 1.0 / a[j + j * m];
      for (int i = j + 1; i < m; ++i) a[i + j * m] = __dmul_rn(a[i + j * m], r);
    }
    for (int k = j + 1; k < m; ++k) for (int i = j + 1; i < m; ++i) a[i + k * m] = __dsub_rn(a[i + k * m], __dmul_rn(a[i + j * m], a[j + k * m]));
  }
  auto lusolve = [&](double *b) {                              // dgetrs 'n'
    for (int j = 0; j < m; ++j) if (piv[j] != j) { const double t = b[j]; b[j] = b[piv[j]]; b[piv[j]] = t; }
    for (int j = 0; j < m; ++j) for (int i = j + 1; i < m; ++i) b[i] = __dsub_rn(b[i], __dmul_rn(b[j], a[i + j * m]));
    for (int j = m - 1; j >= 0; --j) {
      b[j] = b[j] / a[j + j * m];
      for (int i = 0; i < j; ++i) b[i] = __dsub_rn(b[i], __dmul_rn(b[j], a[i + j * m]));
    }
  };
  auto symv = [&](const double *A, const double *xv, double *yv) {   // dsymv 'u', reference BLAS loop order
    for (int i = 0; i < L + 1; ++i) yv[i] = 0.0;
    for (int j = 0; j < L + 1; ++j) {
      const double t1 = xv[j]; double t2 = 0.0;
      for (int i = 0; i < j; ++i) { yv[i] = __dadd_rn(yv[i], __dmul_rn(t1, A[i + j * ldr])); t2 = __dadd_rn(t2, __dmul_rn(A[i + j * ldr], xv[i])); }
      yv[j] = __dadd_rn(__dadd_rn(yv[j], __dmul_rn(t1, A[j + j * ldr])), t2);
    }
  };
  auto dots = [&](const double *xv, const double *yv) { double s = 0.0; for (int i = 0; i < L + 1; ++i) s = __dadd_rn(s, __dmul_rn(xv[i], yv[i])); return s; };
  RW(1, y0) = -1.0;
  for (int i = 2; i <= L; ++i) RW(i, y0) = RW(i, z);
  for (int i = 1; i <= m; ++i) tv[i - 1] = RW(i + 1, y0);
  lusolve(tv);
  for (int i = 1; i <= m; ++i) RW(i + 1, y0) = tv[i - 1];
  RW(L + 1, y0) = 0.0;
  RW(1, yl) = 0.0;
  for (int i = 1; i <= m; ++i) { RW(i + 1, yl) = RW(i + 1, z + L); tv[i - 1] = RW(i + 1, yl); }
  lusolve(tv);
  for (int i = 1; i <= m; ++i) RW(i + 1, yl) = tv[i - 1];
  RW(L + 1, yl) = -1.0;
  symv(&RW(1, z), &RW(1, y0), &RW(1, y));
  double kappa0 = dots(&RW(1, y0), &RW(1, y));
  if (kappa0 <= 0.0) { c->info = HUTI_HALTED; c->done = 1; return; }
  kappa0 = sqrt(kappa0);
  symv(&RW(1, z), &RW(1, yl), &RW(1, y));
  double kappal = dots(&RW(1, yl), &RW(1, y));
  if (kappal <= 0.0) { c->info = HUTI_HALTED; c->done = 1; return; }
  kappal = sqrt(kappal);
  symv(&RW(1, z), &RW(1, y0), &RW(1, y));
  const double varrho = dots(&RW(1, yl), &RW(1, y)) / __dmul_rn(kappa0, kappal);
  const double hatgamma = __dmul_rn(__dmul_rn(varrho / fabs(varrho), fmax(fabs(varrho), 7e-1)), kappa0) / kappal;
  for (int i = 1; i <= L + 1; ++i) RW(i, y0) = __dsub_rn(RW(i, y0), __dmul_rn(hatgamma, RW(i, yl)));
  sc[BL_OMEGA] = RW(L + 1, y0);
  for (int j = 1; j <= L; ++j) sc[BL_GAM + j] = RW(j + 1, y0);
  symv(&RW(1, z), &RW(1, y0), &RW(1, y));
  double rnrm = dots(&RW(1, y0), &RW(1, y));
  if (rnrm < 0.0) { c->info = HUTI_HALTED; c->done = 1; return; }
  rnrm = sqrt(rnrm);
  sc[BL_RNRM] = rnrm;
  // 1050-1068
  double mxx = fmax(sc[BL_MXX], rnrm), mxr = fmax(sc[BL_MXR], rnrm);
  const double rnrm0 = sc[BL_RNRM0];
  const bool xpdt = (rnrm < 1.0e-2 * rnrm0 && rnrm0 < mxx);
  const bool rcmp = ((rnrm < 1.0e-2 * mxr && rnrm0 < mxr) || xpdt);
  if (rcmp) mxr = rnrm;
  if (xpdt) mxx = rnrm;
  sc[BL_MXX] = mxx; sc[BL_MXR] = mxr; sc[BL_XPDT] = xpdt ? 1.0 : 0.0; sc[BL_RCMP] = rcmp ? 1.0 : 0.0;
  c->residual = rnrm / sc[BL_BNRM];
#undef RW
}
// 1016-1032: per element, for j = 1..l in order: u_0 -= g_j u_j ; x += g_j r_(j-1) ; r_0 -= g_j r_j
__global__ void __launch_bounds__(256) k_bl_gamma(int n, const Ctrl *c, const double *sc, double *__restrict__ x, BlVecs v, int l) {
  if (c->done) return;
  double g[BL_MAXL + 1];
  for (int j = 1; j <= l; ++j) g[j] = sc[BL_GAM + j];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double u0 = v.u[0][i], xi = x[i], r0 = v.r[0][i];
    for (int j = 1; j <= l; ++j) {
      u0 = __dsub_rn(u0, __dmul_rn(g[j], v.u[j][i]));
      const double rjm1 = (j == 1) ? r0 : v.r[j - 1][i];       // j = 1: r_0 as it is before this j's own update (x is updated before r_0)
      xi = __dadd_rn(xi, __dmul_rn(g[j], rjm1));
      r0 = __dsub_rn(r0, __dmul_rn(g[j], v.r[j][i]));
    }
    v.u[0][i] = u0; x[i] = xi; v.r[0][i] = r0;
  }
}
// opens / closes the conditional section of the reliable update: kernels in between see Ctrl::done != 0 and return
__global__ void k_bl_section(Ctrl *c, const double *sc, int open) {
  if (open) { if (c->done == 0 && sc[BL_RCMP] == 0.0) c->done = 2; }
  else if (c->done == 2) c->done = 0;
}
// 1060-1068: r = b' - r ; flying restart: x' += t, x = 0, b' = r
__global__ void __launch_bounds__(256) k_bl_rcmp(int n, const Ctrl *c, const double *sc, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ rnew,
                                                  double *__restrict__ bp, double *__restrict__ xp, const double *__restrict__ t) {
  if (c->done) return;
  const bool xpdt = sc[BL_XPDT] != 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double ri = __dsub_rn(bp[i], rnew[i]);
    r[i] = ri;
    if (xpdt) { xp[i] = __dadd_rn(xp[i], t[i]); x[i] = 0.0; bp[i] = ri; }
  }
}
// 1104-1131: end of a round
__global__ void k_bl_roundend(Ctrl *c, double *sc, double Tol, double MaxTol, int MaxRounds) {
  if (c->done) return;
  const double errorind = sc[BL_RNRM] / sc[BL_BNRM];
  c->residual = errorind;
  const int Round = (int)sc[BL_ROUND];
  c->iters = Round;
  if (errorind < Tol) { c->info = HUTI_CONVERGENCE; c->done = 1; return; }
  if (errorind > MaxTol || errorind != errorind) { c->info = HUTI_DIVERGENCE; c->done = 1; return; }
  if (Round + 1 > MaxRounds) { c->info = HUTI_MAXITER; c->done = 1; return; }
  sc[BL_ROUND] = (double)(Round + 1);
}

static HostResult run_bicgstabl_dev(Handle &h, const double *b, double *x, int pc, int MaxRounds, double Tol, double MaxTol, int l) {
  HostResult res;
  const int nw = 3 + 2 * (l + 1);
  Solver S(h, pc, nw + 2);
  const int n = S.n;
  cudaStream_t st = S.st; Ctrl *ctrl = S.ctrl; double *sc = S.sc;
  auto work = [&](int c) { return S.vec[c - 1]; };
  double *t = S.vec[nw], *rnew = S.vec[nw + 1];                // rnew: A M^-1 x of the reliable update (partitioned SpMVs do not look at Ctrl::done)
  const int rr = 1, r = rr + 1, u = r + (l + 1), xp = u + (l + 1), bp = xp + 1;
  { double nx2 = S.dot(x, x); if (nx2 == 0.0) copy_vec(h, n, b, x); }                 // 719
  S.matvec(x, work(r));
  S.lin(b, 1.0, work(r), -1.0);                               // r = b - r
  double bnrm, rnrm0;
  { const double *xs[2] = {b, work(r)}, *ys[2] = {b, work(r)}; double o[2]; S.dots(2, xs, ys, o); bnrm = sqrt(o[0]); rnrm0 = sqrt(o[1]); }
  double errorind = rnrm0 / bnrm;
  if (bnrm != bnrm || rnrm0 != rnrm0 || errorind != errorind) { res.info = HUTI_DIVERGENCE; res.residual = errorind; return res; }
  if (errorind < Tol || errorind > MaxTol) { res.info = errorind < Tol ? HUTI_CONVERGENCE : HUTI_DIVERGENCE; res.residual = errorind; return res; }
  copy_vec(h, n, work(r), work(rr)); copy_vec(h, n, work(r), work(bp));
  copy_vec(h, n, x, work(xp));
  fill_vec(h, n, x, 0.0);
  k_bl_init<<<1, 1, 0, st>>>(sc, bnrm, rnrm0);
  BlVecs V;
  for (int j = 0; j <= BL_MAXL; ++j) { V.r[j] = work(r + std::min(j, l)); V.u[j] = work(u + std::min(j, l)); }
  const double *gx[NRED], *gy[NRED]; int ng = 0;
  for (int i = 1; i <= l + 1; ++i) for (int j = 1; j <= i; ++j) { gx[ng] = work(r + i - 1); gy[ng] = work(r + j - 1); ++ng; }
  Poller poll(h);
  auto pcond = [&](double *dst, double *src) -> double * { return S.precond(dst, src); };
  int Round = 0;
  for (Round = 1; Round <= MaxRounds; ++Round) {
    dot1(h, n, work(rr), work(r), sc + BL_RHO1); reduce_scalars(h, sc + BL_RHO1, 1);          // rho1 of k = 1
    for (int k = 1; k <= l; ++k) {
      k_bl_rho<<<1, 1, 0, st>>>(ctrl, sc, k == 1 ? 1 : 0);
      k_bl_uupd<<<S.blocks, 256, 0, st>>>(n, ctrl, sc, V, k);
      { double *tt = pcond(t, work(u + k - 1));
        SpmvArgs a; a.x = tt; a.y = work(u + k); a.w = work(rr); a.out = sc + BL_SIGMA; a.ctrl = ctrl; spmv_any(h, a, EPI_DOT1); reduce_scalars(h, sc + BL_SIGMA, 1); }
      k_bl_alpha<<<1, 1, 0, st>>>(ctrl, sc);
      k_bl_xr<<<S.blocks, 256, 0, st>>>(n, ctrl, sc, x, V, k, h.red_partials.p, h.red_counters.p);
      { double *tt = pcond(t, work(r + k - 1));
        SpmvArgs a; a.x = tt; a.y = work(r + k); a.w = work(rr); a.out = sc + BL_RHO1; a.ctrl = ctrl; spmv_any(h, a, EPI_DOT1); reduce_scalars(h, sc + BL_RNRM2, 2); }
      k_bl_check<<<1, 1, 0, st>>>(ctrl, sc, Tol);
      h.st_launch += 5;
    }
    dot_batch(h, n, ng, gx, gy, sc + BL_GRAM); reduce_scalars(h, sc + BL_GRAM, ng);
    switch (l) {
      case 2: k_bl_poly<2><<<1, 1, 0, st>>>(ctrl, sc); break;
      case 3: k_bl_poly<3><<<1, 1, 0, st>>>(ctrl, sc); break;
      case 4: k_bl_poly<4><<<1, 1, 0, st>>>(ctrl, sc); break;
      default: k_bl_poly<5><<<1, 1, 0, st>>>(ctrl, sc); break;
    }
    k_bl_gamma<<<S.blocks, 256, 0, st>>>(n, ctrl, sc, x, V, l);
    k_bl_section<<<1, 1, 0, st>>>(ctrl, sc, 1);
    { double *tt = pcond(t, x);
      SpmvArgs a; a.x = tt; a.y = rnew; a.ctrl = ctrl; h.mv_honor_skip = true; spmv_any(h, a, EPI_NONE); h.mv_honor_skip = false;
      k_bl_rcmp<<<S.blocks, 256, 0, st>>>(n, ctrl, sc, x, work(r), rnew, work(bp), work(xp), tt); }
    k_bl_section<<<1, 1, 0, st>>>(ctrl, sc, 0);
    k_bl_roundend<<<1, 1, 0, st>>>(ctrl, sc, Tol, MaxTol, MaxRounds);
    h.st_launch += 6;
    B200_CUDA(cudaGetLastError());
    {                                                          // poll one round behind; only done == 1 ends the solve (2 = section skip)
      int cur = Round & 1, prev = cur ^ 1;
      B200_CUDA(cudaMemcpyAsync(poll.slot[cur], h.ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaEventRecord(poll.ev[cur], st));
      poll.pending[cur] = true;
      if (poll.pending[prev]) {
        B200_CUDA(cudaEventSynchronize(poll.ev[prev]));
        poll.pending[prev] = false;
        if (poll.slot[prev]->done == 1 || poll.slot[prev]->spin_timeout) break;
      }
    }
  }
  read_ctrl(h);
  B200_REQUIRE(h.h_ctrl->done == 1 || h.h_ctrl->spin_timeout, "BiCGStab(l): the device loop ended without a verdict");
  res.iters = std::min(MaxRounds, h.h_ctrl->iters);
  res.residual = h.h_ctrl->residual;
  res.info = h.h_ctrl->info;
  // the kernels of the solve must not see done = 1 any more: the final assembly x = M^-1 x + x' (1156-1166) is unconditional
  B200_CUDA(cudaMemsetAsync(&h.ctrl.p->done, 0, sizeof(int), st));
  copy_vec(h, n, x, t);
  if (pc != 0) { S.precond(x, t); } else h.st_pcond++;
  S.lin(work(xp), 1.0, x, 1.0);
  return res;
}

// IterativeMethods.F90:1260-1458
static HostResult run_gcr(Handle &h, const double *b, double *x, int pc, int Rounds, double MinTol, double MaxTol, int m, int MinIter) {
  HostResult res;
  B200_REQUIRE(m >= 1, "GCR: restart < 1");
  const int nS = std::max(0, m - 1);
  Solver S(h, pc, 3 + 2 * nS);
  const int n = S.n;
  double *R = S.vec[0], *T1 = S.vec[1], *T2 = S.vec[2];
  auto Sc = [&](int j) { return S.vec[3 + (j - 1)]; };
  auto Vc = [&](int j) { return S.vec[3 + nS + (j - 1)]; };
  bool Converged = false, Diverged = false;
  S.matvec(x, R);
  S.lin(b, 1.0, R, -1.0);
  double bnorm, rnorm;
  { const double *xs[2] = {b, R}, *ys[2] = {b, R}; double o[2]; S.dots(2, xs, ys, o); bnorm = sqrt(o[0]); rnorm = sqrt(o[1]); }
  double Residual = rnorm / bnorm;
  Converged = (Residual < MinTol) && (MinIter <= 0);
  Diverged = (Residual > MaxTol) || (Residual != Residual);
  int k = 0;
  if (!(Converged || Diverged)) {
    for (k = 1; k <= Rounds; ++k) {
      int j;
      if (k % m == 0) j = m;
      else {
        j = k % m;
        if (j == 1 && k > 1) { S.matvec(x, R); S.lin(b, 1.0, R, -1.0); }
      }
      double *t1src = S.precond(T1, R);
      S.matvec(t1src, T2);
      if (t1src != T1) copy_vec(h, n, t1src, T1);              // T1 is modified below; R must stay intact
      for (int i = 1; i <= j - 1; ++i) {                       // sequential (classical) Gram-Schmidt, 1338-1364
        double beta = S.dot(Vc(i), T2);
        LinOp ops[2] = {LinOp{Sc(i), T1, -beta, 1.0}, LinOp{Vc(i), T2, -beta, 1.0}};
        axpby_batch(h, n, 2, ops);
      }
      double alpha = S.norm(T2);
      { double ia = 1.0 / alpha; LinOp ops[2] = {LinOp{T1, T1, ia, 0.0}, LinOp{T2, T2, ia, 0.0}}; axpby_batch(h, n, 2, ops); }
      double beta = S.dot(T2, R);
      { LinOp ops[2] = {LinOp{T1, x, beta, 1.0}, LinOp{T2, R, -beta, 1.0}}; axpby_batch(h, n, 2, ops); }
      if (j != m) { copy_vec(h, n, T1, Sc(j)); copy_vec(h, n, T2, Vc(j)); }
      rnorm = S.norm(R);
      Residual = rnorm / bnorm;
      Converged = (Residual < MinTol) && (k >= MinIter);
      // 1427-1431: the reference recomputes the true residual here for an informational message only
      Diverged = (Residual > MaxTol) || (Residual != Residual);
      if (Converged || Diverged) break;
    }
  }
  res.iters = std::min(k, Rounds); res.residual = Residual;
  if (Converged) res.info = HUTI_CONVERGENCE;
  if (Diverged) res.info = HUTI_DIVERGENCE;
  if (!Converged && !Diverged) res.info = HUTI_MAXITER;
  return res;
}

// fhutiter/src/huti_gmres.F90:390-822 huti_dgmressolv: restarted GMRES(m), modified Gram-Schmidt, Givens rotations.
// IterSolver hands GMRES the preconditioner in the LEFT slot (IterSolve.F90:509-525), so every residual here is
// M^-1 (b - A x).  Vectors stay on the device; the (m+1)^2 Hessenberg algebra and huti_dlusolve run on the host
// exactly as in the reference.  HUTI_ITERS counts restart cycles.
static void host_lusolve(int n, std::vector<double> &lu, double *u, const double *v) {   // huti_aux.F90:221-289
  auto LU = [&](int i, int j) -> double & { return lu[(size_t)(i - 1) + (size_t)(j - 1) * n]; };
  for (int i = 2; i <= n; ++i)
    for (int k = 1; k <= i - 1; ++k) {
      LU(i, k) = LU(i, k) / LU(k, k);
      for (int j = k + 1; j <= n; ++j) LU(i, j) = LU(i, j) - LU(i, k) * LU(k, j);
    }
  for (int i = 1; i <= n; ++i) { u[i - 1] = v[i - 1]; for (int k = 1; k <= i - 1; ++k) u[i - 1] = u[i - 1] - LU(i, k) * u[k - 1]; }
  for (int i = n; i >= 1; --i) { for (int k = i + 1; k <= n; ++k) u[i - 1] = u[i - 1] - LU(i, k) * u[k - 1]; u[i - 1] = u[i - 1] / LU(i, i); }
}
static HostResult run_gmres(Handle &h, const double *b, double *x, int pc, int MaxIt, double Tol, double MaxTol, int m, int stopc) {
  HostResult res;
  B200_REQUIRE(m >= 1, "GMRES: restart < 1");
  Solver S(h, pc, 3 + m + 1);
  const int n = S.n;
  double *W = S.vec[0], *R = S.vec[1], *T1V = S.vec[2];
  auto V = [&](int i) { return S.vec[3 + (i - 1)]; };
  std::vector<double> H((size_t)(m + 1) * (m + 1), 0.0), HLU, CS(m + 2, 0.0), SN(m + 2, 0.0), Y(m + 1, 0.0), Sv(m + 2, 0.0);
  auto Hh = [&](int i, int j) -> double & { return H[(size_t)(i - 1) + (size_t)(j - 1) * (m + 1)]; };
  const double bnrm = S.norm(b);
  const double rhsnorm = (stopc == 1 || stopc == 3) ? bnrm : 1.0;
  // M^-1 (b - A x) into R (or the vector the preconditioner aliased it to)
  auto prec_residual = [&]() -> double * {
    S.matvec(x, R);
    copy_vec(h, n, b, T1V); S.lin(R, -1.0, T1V, 1.0);            // T1V = B - R
    return S.precond(R, T1V);
  };
  h.st_matvec++; h.st_pcond++;                                    // the reference's initial residual (475-486) is recomputed at 300: not repeated here
  int iter_count = 1; double residual = 0.0;
  auto update_x = [&](int k) {                                    // X = X + V(:,1:k) Y
    HLU.assign((size_t)k * k, 0.0);
    for (int c = 1, j = 0; c <= k; ++c) for (int l = 1; l <= k; ++l) HLU[j++] = Hh(l, c);
    host_lusolve(k, HLU, Y.data(), Sv.data());
    for (int c = 1; c <= k; ++c) S.lin(V(c), Y[c - 1], x, 1.0);
  };
  for (;;) {
    double *rs = prec_residual();
    const double alpha = S.norm(rs);
    if (alpha == 0) { res.info = 40; break; }                     // HUTI_GMRES_ALPHA
    div_scalar(h, n, rs, V(1), alpha);                            // V(:,1) = R / alpha, huti_gmres.F90:189
    std::fill(Sv.begin(), Sv.end(), 0.0); Sv[0] = alpha;          // S = alpha * e1
    bool early = false, broke = false;
    for (int i = 1; i <= m; ++i) {
      S.matvec(V(i), T1V);
      double *ws = S.precond(W, T1V);
      if (ws != W) copy_vec(h, n, ws, W);
      for (int k = 1; k <= i; ++k) {
        Hh(k, i) = S.dot(W, V(k));
        S.lin(V(k), -Hh(k, i), W, 1.0);
      }
      const double beta = S.norm(W);
      if (beta == 0) { res.info = 41; broke = true; break; }      // HUTI_GMRES_BETA
      Hh(i + 1, i) = beta;
      div_scalar(h, n, W, V(i + 1), beta);
      for (int k = 1; k <= i - 1; ++k) {
        const double temp = CS[k] * Hh(k, i) + SN[k] * Hh(k + 1, i);
        Hh(k + 1, i) = -1 * SN[k] * Hh(k, i) + CS[k] * Hh(k + 1, i);
        Hh(k, i) = temp;
      }
      if (Hh(i + 1, i) == 0) { CS[i] = 1; SN[i] = 0; }
      else if (fabs(Hh(i + 1, i)) > fabs(Hh(i, i))) { const double t2 = Hh(i, i) / Hh(i + 1, i); SN[i] = 1 / sqrt(1 + (t2 * t2)); CS[i] = t2 * SN[i]; }
      else { const double t2 = Hh(i + 1, i) / Hh(i, i); CS[i] = 1 / sqrt(1 + (t2 * t2)); SN[i] = t2 * CS[i]; }
      const double temp = CS[i] * Sv[i - 1];
      Sv[i] = -1 * SN[i] * Sv[i - 1];
      Sv[i - 1] = temp;
      Hh(i, i) = (CS[i] * Hh(i, i)) + (SN[i] * Hh(i + 1, i));
      Hh(i + 1, i) = 0;
      const double error = fabs(Sv[i]) / bnrm;
      if ((float)error < Tol) { update_x(i); early = true; break; }   // 628: REAL(error)
    }
    if (broke) break;
    if (!early) update_x(m);
    if (stopc == 2 || stopc == 3) {                                // pseudo-residual criteria (703-732)
      S.matvec(x, R); S.lin(b, -1.0, R, 1.0);
      residual = S.norm(S.precond(T1V, R)) / rhsnorm;
    } else {
      residual = S.norm(prec_residual()) / rhsnorm;
    }
    if (residual < Tol) { res.info = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > MaxTol) { res.info = HUTI_DIVERGENCE; break; }
    iter_count = iter_count + 1;
    if (iter_count > MaxIt) { res.info = HUTI_MAXITER; break; }
  }
  res.iters = iter_count; res.residual = residual;
  return res;
}

// fhutiter/src/huti_cgs.F90:283-470 huti_dcgssolv (right-oriented preconditioning; the dummy left preconditioner is a copy).
static HostResult run_cgs(Handle &h, const double *b, double *x, int pc, int MaxIt, double Tol, double MaxTol, int stopc) {
  HostResult res;
  Solver S(h, pc, 7);
  const int n = S.n;
  double *RTLD = S.vec[0], *P = S.vec[1], *Q = S.vec[2], *U = S.vec[3], *T1V = S.vec[4], *T2V = S.vec[5], *R = S.vec[6];
  const double rhsnorm = (stopc == 1 || stopc == 3) ? S.norm(b) : 1.0;
  S.matvec(x, R); S.lin(b, 1.0, R, -1.0);                          // R = B - A X
  copy_vec(h, n, R, RTLD);
  double rho = 0, oldrho = 0, alpha = 0, beta = 0, residual = 0;
  int iter_count = 1;
  for (;;) {
    rho = S.dot(RTLD, R);
    if (rho == 0) { res.info = 25; break; }                        // HUTI_CGS_RHO
    if (iter_count == 1) { copy_vec(h, n, R, U); copy_vec(h, n, U, P); }
    else {
      beta = rho / oldrho;
      copy_vec(h, n, R, U); S.lin(Q, beta, U, 1.0);                // U = R + beta Q
      S.lin(U, 1.0, P, beta * beta); S.lin(Q, beta, P, 1.0);       // P = U + beta Q + beta^2 P
    }
    double *t1 = S.precond(T1V, P);                                // pcondl = copy, pcondr = M^-1
    S.matvec(t1, T2V);
    alpha = rho / S.dot(RTLD, T2V);
    copy_vec(h, n, U, Q); S.lin(T2V, -alpha, Q, 1.0);              // Q = U - alpha T2V
    S.lin(Q, 1.0, U, 1.0);                                         // U := U + Q  (T2V = U + Q; pcondl(U, T2V) copies it into U)
    t1 = S.precond(T1V, U);
    S.lin(t1, alpha, x, 1.0);
    S.matvec(t1, T2V);
    S.lin(T2V, -alpha, R, 1.0);
    if (stopc == 2 || stopc == 3) residual = S.norm(R) / rhsnorm;
    else { S.matvec(x, T1V); S.lin(b, -1.0, T1V, 1.0); residual = S.norm(T1V) / rhsnorm; }
    if (residual < Tol) { res.info = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > MaxTol) { res.info = HUTI_DIVERGENCE; break; }
    oldrho = rho;
    iter_count = iter_count + 1;
    if (iter_count > MaxIt) { res.info = HUTI_MAXITER; break; }
  }
  res.iters = iter_count; res.residual = residual;
  return res;
}

// fhutiter/src/huti_tfqmr.F90:455-803 huti_dtfqmrsolv; preconditioner in the left slot as IterSolver passes it (IterSolve.F90:509-525).
static HostResult run_tfqmr(Handle &h, const double *b, double *x, int pc, int MaxIt, double Tol, double MaxTol, int stopc) {
  HostResult res;
  Solver S(h, pc, 10);
  const int n = S.n;
  double *V = S.vec[0], *Y = S.vec[1], *YNEW = S.vec[2], *RTLD = S.vec[3], *T1V = S.vec[4], *T2V = S.vec[5], *W = S.vec[6], *D = S.vec[7],
         *R = S.vec[8], *TRV = S.vec[9];
  const double rhsnorm = (stopc == 1 || stopc == 3) ? S.norm(b) : 1.0;
  double rho = 0, oldrho = 0, eta = 0, tau = 0, gamma = 0, oldgamma = 0, alpha = 0, beta = 0, c = 0, residual = 0;
  int iter_count = 1;
  // dst = M^-1 A src (through tmp): pcondr is the dummy, pcondl the preconditioner
  auto apply = [&](const double *src, double *tmp, double *dst) { S.matvec(src, tmp); double *r = S.precond(dst, tmp); if (r != dst) copy_vec(h, n, r, dst); };
  auto check = [&]() {
    if (stopc == 2 || stopc == 3) { S.matvec(x, R); S.lin(b, -1.0, R, 1.0); residual = S.norm(S.precond(TRV, R)) / rhsnorm; }
    else { S.matvec(x, R); copy_vec(h, n, R, TRV); S.lin(b, -1.0, TRV, 1.0); residual = S.norm(S.precond(R, TRV)) / rhsnorm; }
  };
  auto half = [&](const double *yv, const double *av) {
    S.lin(av, -alpha, W, 1.0);
    gamma = S.norm(W) / tau;
    c = 1 / sqrt(1 + gamma * gamma);
    tau = tau * gamma * c;
    S.lin(yv, 1.0, D, (oldgamma * oldgamma * eta) / alpha);        // D = y + f D
    eta = c * c * alpha;
    S.lin(D, eta, x, 1.0);
    oldgamma = gamma;
  };
  S.matvec(x, R); copy_vec(h, n, b, D); S.lin(R, -1.0, D, 1.0);    // D = B - A X
  { double *r = S.precond(R, D); if (r != R) copy_vec(h, n, r, R); }
  copy_vec(h, n, R, Y); copy_vec(h, n, R, W);
  apply(Y, D, V);
  copy_vec(h, n, V, T2V);
  fill_vec(h, n, D, 0.0);
  tau = S.norm(R);
  copy_vec(h, n, R, RTLD);
  oldrho = S.dot(RTLD, R);
  if (oldrho == 0) res.info = 30;                                    // HUTI_TFQMR_RHO
  else for (;;) {
    alpha = oldrho / S.dot(RTLD, V);
    copy_vec(h, n, Y, YNEW); S.lin(V, -alpha, YNEW, 1.0);
    half(Y, T2V);
    check();
    if (residual < Tol) { res.info = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > MaxTol) { res.info = HUTI_DIVERGENCE; break; }
    apply(YNEW, R, T1V);
    half(YNEW, T1V);
    check();
    if (residual < Tol) { res.info = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > MaxTol) { res.info = HUTI_DIVERGENCE; break; }
    rho = S.dot(RTLD, W);
    beta = rho / oldrho;
    S.lin(W, 1.0, YNEW, beta);                                       // YNEW = W + beta YNEW
    apply(YNEW, R, T2V);
    S.lin(T2V, 1.0, V, beta * beta); S.lin(T1V, beta, V, 1.0);      // V = T2V + beta T1V + beta^2 V
    copy_vec(h, n, YNEW, Y);
    oldrho = rho;
    iter_count = iter_count + 1;
    if (iter_count > MaxIt) { res.info = HUTI_MAXITER; break; }
  }
  res.iters = iter_count; res.residual = residual;
  return res;
}

// fhutiter/src/huti_bicgstab_2.F90:339-578 huti_dbicgstab_2solv; preconditioner in the left slot (IterSolve.F90:509-525).
static HostResult run_bicgstab2(Handle &h, const double *b, double *x, int pc, int MaxIt, double Tol, double MaxTol, int stopc) {
  HostResult res;
  Solver S_(h, pc, 8);
  const int n = S_.n;
  double *RTLD = S_.vec[0], *U = S_.vec[1], *T1V = S_.vec[2], *V = S_.vec[3], *S = S_.vec[4], *W = S_.vec[5], *T = S_.vec[6], *R = S_.vec[7];
  const double rhsnorm = (stopc == 1 || stopc == 3) ? S_.norm(b) : 1.0;
  double rho = 0, oldrho = 1, alpha = 0, beta = 0, omega1 = 0, omega2 = 1, residual = 0;
  int iter_count = 1;
  auto apply = [&](double *dst, const double *src) { S_.matvec(src, T1V); double *r = S_.precond(dst, T1V); if (r != dst) copy_vec(h, n, r, dst); };
  S_.matvec(x, R); copy_vec(h, n, b, U); S_.lin(R, -1.0, U, 1.0);
  { double *r = S_.precond(R, U); if (r != R) copy_vec(h, n, r, R); }
  copy_vec(h, n, R, RTLD); fill_vec(h, n, U, 0.0);
  for (;;) {
    oldrho = -omega2 * oldrho;
    rho = S_.dot(RTLD, R);
    if (rho == 0) { res.info = 45; break; }                        // HUTI_BICGSTAB_2_RHO
    beta = (rho * alpha) / oldrho; oldrho = rho;
    S_.lin(R, 1.0, U, -beta);                                      // U = R - beta U
    apply(V, U);
    alpha = oldrho / S_.dot(RTLD, V);
    S_.lin(V, -alpha, R, 1.0);
    apply(S, R);
    S_.lin(U, alpha, x, 1.0);
    rho = S_.dot(RTLD, S);
    if (rho == 0) { res.info = 45; break; }
    beta = (rho * alpha) / oldrho; oldrho = rho;
    S_.lin(S, 1.0, V, -beta);                                      // V = S - beta V
    apply(W, V);
    alpha = oldrho / S_.dot(RTLD, W);
    S_.lin(R, 1.0, U, -beta);                                      // U = R - beta U
    S_.lin(V, -alpha, R, 1.0);
    S_.lin(W, -alpha, S, 1.0);
    apply(T, S);
    double o[5];
    { const double *xs[5] = {R, S, S, T, R}, *ys[5] = {S, S, T, T, T}; S_.dots(5, xs, ys, o); }
    double myy = o[1], delta = o[2], tau = o[3];
    omega1 = o[0]; omega2 = o[4];
    tau = tau - (delta * delta) / myy;
    omega2 = (omega2 - (delta * omega1) / myy) / tau;
    omega1 = (omega1 - delta * omega2) / myy;
    { LinOp ops[3] = {LinOp{R, x, omega1, 1.0}, LinOp{S, x, omega2, 1.0}, LinOp{U, x, alpha, 1.0}}; for (auto &op : ops) axpby_batch(h, n, 1, &op); }
    S_.lin(S, -omega1, R, 1.0); S_.lin(T, -omega2, R, 1.0);
    if (stopc == 2 || stopc == 3) residual = S_.norm(R) / rhsnorm;
    else {
      S_.matvec(x, T1V); S_.lin(b, -1.0, T1V, 1.0);
      residual = S_.norm(S_.precond(S, T1V)) / rhsnorm;
    }
    if (residual < Tol) { res.info = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > MaxTol) { res.info = HUTI_DIVERGENCE; break; }
    S_.lin(V, -omega1, U, 1.0); S_.lin(W, -omega2, U, 1.0);
    iter_count = iter_count + 1;
    if (iter_count > MaxIt) { res.info = HUTI_MAXITER; break; }
  }
  res.iters = iter_count; res.residual = residual;
  return res;
}

// IterativeMethods.F90:336-391 Jacobi, 444-519 Richardson: x += r / a_ii, resp. x = b / m (first round), x += r / m with m = the row
// sums (summed left to right by one thread per row, as the reference loop does).
__global__ void k_row_sums(int n, const int *__restrict__ rows, const double *__restrict__ vals, double *__restrict__ m) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int j = rows[i]; j < rows[i + 1]; ++j) s = __dadd_rn(s, vals[j]);
    m[i] = s;
  }
}
__global__ void k_stationary_update(int n, const double *__restrict__ d, const double *__restrict__ r, const double *__restrict__ b, double *x, int first) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    x[i] = first ? __ddiv_rn(b[i], d[i]) : __dadd_rn(x[i], __ddiv_rn(r[i], d[i]));
}
static HostResult run_stationary(Handle &h, const double *b, double *x, int Rounds, double MinTol, double MaxTol, bool richardson) {
  HostResult res;
  B200_REQUIRE(h.nranks == 1, "Jacobi / Richardson iterations are implemented for single-rank handles only");
  Solver S(h, 0, 2);
  const int n = S.n;
  double *R = S.vec[0], *M = S.vec[1];
  S.matvec(x, R); S.lin(b, 1.0, R, -1.0);
  double o[2];
  { const double *xs[2] = {b, R}, *ys[2] = {b, R}; S.dots(2, xs, ys, o); }
  const double bnorm = sqrt(o[0]);
  double Residual = sqrt(o[1]) / bnorm;
  bool Converged = Residual < MinTol, Diverged = (Residual > MaxTol) || (Residual != Residual);
  int k = 0;
  if (!(Converged || Diverged)) {
    if (richardson) k_row_sums<<<S.blocks, 256, 0, S.st>>>(n, h.d_rows.p, h.d_vals.p, M);
    const double *d = richardson ? M : h.d_dvals.p;
    for (k = 1; k <= Rounds; ++k) {
      k_stationary_update<<<S.blocks, 256, 0, S.st>>>(n, d, R, b, x, (richardson && k == 1) ? 1 : 0);
      h.st_launch++;
      S.matvec(x, R); S.lin(b, 1.0, R, -1.0);
      Residual = S.norm(R) / bnorm;
      Converged = Residual < MinTol; Diverged = (Residual > MaxTol) || (Residual != Residual);
      if (Converged || Diverged) break;
    }
  }
  res.iters = std::min(k, Rounds); res.residual = Residual;
  res.info = Converged ? HUTI_CONVERGENCE : (Diverged ? HUTI_DIVERGENCE : HUTI_MAXITER);
  return res;
}

// IterativeMethods.F90:219-283 SGS: rounds of one forward + one backward Gauss-Seidel sweep, the true residual after each round.
static HostResult run_sgs(Handle &h, const double *b, double *x, int Rounds, double MinTol, double MaxTol, double omega) {
  HostResult res;
  B200_REQUIRE(h.nranks == 1, "SGS is implemented for single-rank handles only");
  Solver S(h, 0, 3);
  const int n = S.n;
  double *R = S.vec[0], *T1 = S.vec[1], *T2 = S.vec[2];
  S.matvec(x, R); S.lin(b, 1.0, R, -1.0);
  double o[2];
  { const double *xs[2] = {b, R}, *ys[2] = {b, R}; S.dots(2, xs, ys, o); }
  const double bnorm = sqrt(o[0]);
  double Residual = sqrt(o[1]) / bnorm;
  bool Converged = Residual < MinTol, Diverged = (Residual > MaxTol) || (Residual != Residual);
  int k = 0;
  if (!(Converged || Diverged)) {
    for (k = 1; k <= Rounds; ++k) {
      sgs_sweeps(h, b, x, T1, T2, omega);
      S.matvec(x, R); S.lin(b, 1.0, R, -1.0);
      Residual = S.norm(R) / bnorm;
      Converged = Residual < MinTol; Diverged = (Residual > MaxTol) || (Residual != Residual);
      if (Converged || Diverged) break;
    }
  }
  res.iters = std::min(k, Rounds); res.residual = Residual;
  res.info = Converged ? HUTI_CONVERGENCE : (Diverged ? HUTI_DIVERGENCE : HUTI_MAXITER);
  return res;
}

// counter-based uniform [0,1) generator for the IDR(s) shadow space when the caller passes none
__global__ void k_shadow_space(long long n, double *P, unsigned long long seed) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; z ^= z >> 31;
    P[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0);
  }
}

// IterativeMethods.F90:1579-1913
static HostResult run_idrs(Handle &h, const double *b, double *x, int pc, int MaxRounds, double Tol, double MaxTol, int s,
                           bool Smoothing, const double *d_P, long long goffset_seed, const RobustPar &Rb) {
  HostResult res;
  B200_REQUIRE(s >= 1, "IDR(s): s < 1");
  Solver S(h, pc, 3 * s + 5 + (Rb.on ? 1 : 0));
  double *Bestx = Rb.on ? S.vec[3 * s + 5] : nullptr;          // 1535-1544
  double BestNorm = sqrt(DBL_MAX);
  int BadIterCount = 0;
  const int n = S.n;
  auto P = [&](int j) { return S.vec[j - 1]; };
  auto G = [&](int j) { return S.vec[s + j - 1]; };
  auto U = [&](int j) { return S.vec[2 * s + j - 1]; };
  double *r = S.vec[3 * s], *v = S.vec[3 * s + 1], *t = S.vec[3 * s + 2], *r_s = S.vec[3 * s + 3], *x_s = S.vec[3 * s + 4];
  std::vector<double> M((size_t)s * s, 0.0), f(s), mu(s), alpha(s), beta(s), gamma(s);
  auto Mm = [&](int i, int j) -> double & { return M[(i - 1) + (size_t)(j - 1) * s]; };
  bool Converged = false, Diverged = false;
  double om = 1.0, kappa = 0.7, errorind, normr;
  int iter = 0, jj = 0;
  double normb = S.norm(b);
  S.matvec(x, t);
  copy_vec(h, n, b, r); S.lin(t, -1.0, r, 1.0);               // r = b - t
  normr = S.norm(r);
  errorind = normr / normb;
  Converged = errorind < Tol;
  Diverged = (errorind > MaxTol) || (errorind != errorind);
  if (Converged || Diverged) { res.info = Converged ? HUTI_CONVERGENCE : HUTI_DIVERGENCE; res.residual = errorind; return res; }
  if (Smoothing) { copy_vec(h, n, x, x_s); copy_vec(h, n, r, r_s); }
  if (d_P) { for (int j = 1; j <= s; ++j) copy_vec(h, n, d_P + (size_t)(j - 1) * n, P(j)); }
  else { for (int j = 1; j <= s; ++j) { k_shadow_space<<<S.blocks, 256, 0, S.st>>>(n, P(j), 314159265ULL + 7919ULL * j + (unsigned long long)goffset_seed * 104729ULL); } }
  for (int j = 1; j <= s; ++j) {                               // Gram-Schmidt on P, 1655-1661
    for (int k = 1; k <= j - 1; ++k) { alpha[k - 1] = S.dot(P(k), P(j)); S.lin(P(k), -alpha[k - 1], P(j), 1.0); }
    double nr = S.norm(P(j));
    div_scalar(h, n, P(j), P(j), nr);                          // P(:,j) = P(:,j)/norm: a true division (1659)
  }
  while (!Converged && !Diverged) {
    {                                                          // f = P' r, one batched pass (1682-1684)
      std::vector<const double *> xs(s), ys(s);
      for (int k = 1; k <= s; ++k) { xs[k - 1] = P(k); ys[k - 1] = r; }
      for (int d0 = 0; d0 < s; d0 += NRED) S.dots(std::min(NRED, s - d0), xs.data() + d0, ys.data() + d0, f.data() + d0);
    }
    for (int k = 1; k <= s; ++k) {
      copy_vec(h, n, r, v);
      if (jj > 0) {
        for (int i = k; i <= s; ++i) {                         // 1696-1703
          gamma[i - 1] = f[i - 1];
          for (int j = k; j <= i - 1; ++j) gamma[i - 1] -= Mm(i, j) * gamma[j - 1];
          gamma[i - 1] = gamma[i - 1] / Mm(i, i);
          S.lin(G(i), -gamma[i - 1], v, 1.0);
        }
        double *tt = S.precond(t, v);
        if (tt != t) copy_vec(h, n, tt, t);
        S.lin(t, om, t, 0.0);                                  // t = om*t
        for (int i = k; i <= s; ++i) S.lin(U(i), gamma[i - 1], t, 1.0);
        copy_vec(h, n, t, U(k));
      } else {
        double *tt = S.precond(U(k), v);
        if (tt != U(k)) copy_vec(h, n, tt, U(k));
      }
      S.matvec(U(k), G(k));
      {                                                        // mu = P' G_k (1724-1726)
        std::vector<const double *> xs(s), ys(s);
        for (int i = 1; i <= s; ++i) { xs[i - 1] = P(i); ys[i - 1] = G(k); }
        for (int d0 = 0; d0 < s; d0 += NRED) S.dots(std::min(NRED, s - d0), xs.data() + d0, ys.data() + d0, mu.data() + d0);
      }
      for (int i = 1; i <= k - 1; ++i) {                       // 1727-1736
        alpha[i - 1] = mu[i - 1];
        for (int j = 1; j <= i - 1; ++j) alpha[i - 1] -= Mm(i, j) * alpha[j - 1];
        alpha[i - 1] = alpha[i - 1] / Mm(i, i);
        LinOp ops[2] = {LinOp{G(i), G(k), -alpha[i - 1], 1.0}, LinOp{U(i), U(k), -alpha[i - 1], 1.0}};
        axpby_batch(h, n, 2, ops);
        for (int q = k; q <= s; ++q) mu[q - 1] -= Mm(q, i) * alpha[i - 1];
      }
      for (int q = k; q <= s; ++q) Mm(q, k) = mu[q - 1];
      if (fabs(Mm(k, k)) <= DBL_MIN) { Diverged = true; break; }
      beta[k - 1] = f[k - 1] / Mm(k, k);
      { LinOp ops[2] = {LinOp{G(k), r, -beta[k - 1], 1.0}, LinOp{U(k), x, beta[k - 1], 1.0}}; axpby_batch(h, n, 2, ops); }
      if (k < s) for (int q = k + 1; q <= s; ++q) f[q - 1] -= beta[k - 1] * Mm(q, k);
      if (Smoothing) {
        copy_vec(h, n, r_s, t); S.lin(r, -1.0, t, 1.0);        // t = r_s - r
        const double *xs[2] = {t, t}, *ys[2] = {r_s, t}; double o[2]; S.dots(2, xs, ys, o);
        double theta = o[0] / o[1];
        S.lin(t, -theta, r_s, 1.0);
        // x_s = x_s - theta*(x_s - x)
        copy_vec(h, n, x_s, t); S.lin(x, -1.0, t, 1.0); S.lin(t, -theta, x_s, 1.0);
      }
      iter++;
      normr = Smoothing ? S.norm(r_s) : S.norm(r);
      errorind = normr / normb;
      Converged = errorind < Tol;
      Diverged = (errorind > MaxTol) || (errorind != errorind);
      if (Converged || Diverged) break;
      if (iter == MaxRounds) break;
    }
    if (Converged || Diverged) break;
    if (iter == MaxRounds) break;
    jj++;
    double *vv = S.precond(v, r);
    S.matvec(vv, t);
    double nr, nt, tr;
    { const double *xs[3] = {r, t, t}, *ys[3] = {r, t, r}; double o[3]; S.dots(3, xs, ys, o); nr = sqrt(o[0]); nt = sqrt(o[1]); tr = o[2]; }
    double rho = fabs(tr / (nt * nr));
    om = tr / (nt * nt);
    if (rho < kappa) om = om * kappa / rho;
    if (fabs(om) <= DBL_EPSILON) { Diverged = true; break; }
    // x first: without a preconditioner vv aliases r (v = r is elided), and r changes in the second update
    { LinOp ops[2] = {LinOp{vv, x, om, 1.0}, LinOp{t, r, -om, 1.0}}; axpby_batch(h, n, 2, ops); }
    if (Smoothing) {
      copy_vec(h, n, r_s, t); S.lin(r, -1.0, t, 1.0);
      const double *xs[2] = {t, t}, *ys[2] = {r_s, t}; double o[2]; S.dots(2, xs, ys, o);
      double theta = o[0] / o[1];
      S.lin(t, -theta, r_s, 1.0);
      copy_vec(h, n, x_s, t); S.lin(x, -1.0, t, 1.0); S.lin(t, -theta, x_s, 1.0);
    }
    iter++;
    normr = Smoothing ? S.norm(r_s) : S.norm(r);
    errorind = normr / normb;
    if (Rb.on) {                                               // 1862-1884
      if (errorind < Rb.Step * BestNorm) { BestNorm = errorind; copy_vec(h, n, Smoothing ? x_s : x, Bestx); BadIterCount = 0; }
      else BadIterCount = BadIterCount + 1;
      if (BestNorm < Rb.Tol && (errorind > Rb.MaxTol || BadIterCount > Rb.MaxBadIter)) break;
    }
    Converged = errorind < Tol;
    Diverged = (errorind > MaxTol) || (errorind != errorind);
    if (iter == MaxRounds) break;
  }
  if (Smoothing) copy_vec(h, n, x_s, x);
  if (Rb.on) {                                                 // 1892-1898
    if (BestNorm < Rb.Tol) Converged = true;
    if (BestNorm < errorind) copy_vec(h, n, Bestx, x);
  }
  res.iters = iter; res.residual = errorind;
  if (Converged) res.info = HUTI_CONVERGENCE;
  if (Diverged) res.info = HUTI_DIVERGENCE;
  if (!Converged && !Diverged) res.info = HUTI_MAXITER;
  return res;
}

// ---------------------------------------------------------------------------------------------
// b, x (and P) are device pointers.  Implements the part of IterSolver between the parameter
// parsing and the error mapping (fem/src/IterSolve.F90:470-471, 913, 964-1005).
void solve_device(Handle &h, const double *d_b, double *d_x, int *ipar, double *dpar, int method, int pc, const double *d_P) {
  B200_REQUIRE(h.have_vals, "b200_solve before b200_set_values");
  B200_REQUIRE(method >= 1 && method <= 12, "unknown iterative method");
  B200_REQUIRE(pc >= 0 && pc <= 2, "unknown preconditioner");
  const int n = h.n;
  cudaStream_t st = h.stream;
  const int stopc = IPAR(12);
  B200_REQUIRE(stopc >= 0 && stopc <= 3, "stopping criterion not supported on the device (HUTI_STOPC must be 0..3)");
  if (method >= 3) B200_REQUIRE(stopc != 10, "user stopping criterion not supported");
  long long launches0 = h.st_launch;
  h.st_matvec = 0; h.st_pcond = 0; h.st_d2h = 0;
  if (pc == 2 && !h.ilu_valid) ilu0_factor(h);             // none for the current values yet
  B200_CUDA(cudaEventRecord(h.ev0, st));
  k_init_ctrl<<<1, 1, 0, st>>>(h.ctrl.p, h.scal.p, DPAR(1), DPAR(2), IPAR(10), IPAR(11), stopc);
  h.st_launch++;
  if (method == B200_M_BICGSTAB || method == B200_M_BICGSTABL || method == B200_M_BICGSTAB2) fill_if_all_zero(h, n, d_x, 1.0e-8);   // IterSolve.F90:470-471
  HostResult hr;
  if (n == 0) { IPAR(30) = HUTI_CONVERGENCE; IPAR(31) = 0; return; }
  RobustPar rb;                                               // huti_fdefs.h:132-135, 153-155
  rb.on = IPAR(26) == 1; rb.MaxBadIter = IPAR(27); rb.Start = IPAR(29); rb.Tol = DPAR(3); rb.Step = DPAR(4); rb.MaxTol = DPAR(5);
  switch (method) {
    case B200_M_CG: run_cg(h, d_b, d_x, pc, IPAR(10), stopc); break;
    case B200_M_BICGSTAB: run_bicgstab(h, d_b, d_x, pc, IPAR(10), stopc); break;
    case B200_M_BICGSTABL:
      if (!rb.on && IPAR(16) >= 2 && IPAR(16) <= BL_MAXL && !h.bl_host) hr = run_bicgstabl_dev(h, d_b, d_x, pc, IPAR(10), DPAR(1), DPAR(2), IPAR(16));
      else hr = run_bicgstabl(h, d_b, d_x, pc, IPAR(10), DPAR(1), DPAR(2), IPAR(16), rb);
      break;
    case B200_M_GCR: hr = run_gcr(h, d_b, d_x, pc, IPAR(10), DPAR(1), DPAR(2), IPAR(17), IPAR(11)); break;
    case B200_M_IDRS: hr = run_idrs(h, d_b, d_x, pc, IPAR(10), DPAR(1), DPAR(2), IPAR(18), IPAR(28) == 1, d_P, h.rank, rb); break;
    case B200_M_GMRES: hr = run_gmres(h, d_b, d_x, pc, IPAR(10), DPAR(1), DPAR(2), IPAR(15), stopc); break;
    case B200_M_CGS: hr = run_cgs(h, d_b, d_x, pc, IPAR(10), DPAR(1), DPAR(2), stopc); break;
    case B200_M_TFQMR: hr = run_tfqmr(h, d_b, d_x, pc, IPAR(10), DPAR(1), DPAR(2), stopc); break;
    case B200_M_BICGSTAB2: hr = run_bicgstab2(h, d_b, d_x, pc, IPAR(10), DPAR(1), DPAR(2), stopc); break;
    case B200_M_JACOBI: hr = run_stationary(h, d_b, d_x, IPAR(10), DPAR(1), DPAR(2), false); break;
    case B200_M_RICHARDSON: hr = run_stationary(h, d_b, d_x, IPAR(10), DPAR(1), DPAR(2), true); break;
    case B200_M_SGS: hr = run_sgs(h, d_b, d_x, IPAR(10), DPAR(1), DPAR(2), DPAR(3)); break;
    default: B200_REQUIRE(false, "unknown iterative method code");
  }
  B200_CUDA(cudaEventRecord(h.ev_end, st));
  B200_CUDA(cudaStreamSynchronize(st));
  float ms = 0; B200_CUDA(cudaEventElapsedTime(&ms, h.ev0, h.ev_end));
  h.st_solve_ms = ms;
  if (method <= 2) {
    IPAR(30) = h.h_ctrl->info; IPAR(31) = h.h_ctrl->iters; h.st_resid = h.h_ctrl->residual;
    // callback counts of the reference algorithm for the iterations actually performed (kernels of the
    // iteration queued behind the stopping test return immediately and are not counted)
    const long long performed = std::min(IPAR(31), IPAR(10));
    const bool true_resid = (stopc == 0 || stopc == 1);
    if (method == B200_M_CG) { h.st_matvec = 1 + (true_resid ? 2 : 1) * performed; h.st_pcond = performed; }
    else { h.st_matvec = 1 + (true_resid ? 3 : 2) * performed; h.st_pcond = 2 * performed; }
    if (h.h_ctrl->spin_timeout) { IPAR(30) = HUTI_HALTED; fprintf(stderr, "[elmer_b200] a device-side dependency wait timed out (triangular solve wavefront or halo exchange flags)\n"); }
  } else {
    read_ctrl(h);
    IPAR(30) = hr.info; IPAR(31) = hr.iters; h.st_resid = hr.residual;
    if (h.h_ctrl->spin_timeout) { IPAR(30) = HUTI_HALTED; fprintf(stderr, "[elmer_b200] a device-side dependency wait timed out (triangular solve wavefront or halo exchange flags)\n"); }
  }
  dpar[9] = h.st_resid;
  h.st_iters = IPAR(31);
  h.st_launch_last = h.st_launch - launches0;
}

}  // namespace b200

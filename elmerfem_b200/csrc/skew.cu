// Skewed-lane wavefront triangular solves for the ILU(0) factor of a structured-grid stencil -- CRS_LUSolve, fem/src/CRSMatrix.F90:4590-4663.
// OPT-IN (B200_TRI_MODE=2) and EXPERIMENTAL: written at the end of round 1 from the design in DESIGN.md section 7; the schedule, the
// operand routing and the stream layout are verified on the CPU (profiles/tools/skewed_lane_proto.cpp, tests/test_skew_plan.py run the
// same skewgeom.h code), but this kernel has not yet run on a GPU.  The default (level-scheduled) kernels are untouched.
//
// Why: the level kernel pays one L2 hand-off (~0.9 us) per dependency level, 2 x 1401 of them on the 200^3 heat problem.  Here one warp
// owns a STRIP of <= 32 consecutive lines of one grid plane; lane j solves row A = t - 2j of its line at step t.  With that skew every
// in-plane operand is already in the warp when it is needed:
//   (A-1, B)            the lane's own result of step t-1                     -> register
//   (A-1 | A | A+1, B-1) lane j-1's results of steps t-3 | t-2 | t-1          -> shfl.up of a 3-deep history
// Only line B0-1 (neighbouring strip, lane 0) and the nine operands of the previous plane come from the result vector in L2, which is
// pre-filled with the NaN sentinel and polled exactly as the level kernel does.  Tasks (plane, strip) are taken in increasing order by a
// co-resident grid, every dependency points to a lower task: no deadlock for any grid shape.
// Arithmetic: the reference's operations in the reference's order (entries in ascending column order, separate multiply / subtract
// roundings, inverse diagonal last); pad entries (neighbours outside the grid) are 0 x a finite register.  Results are bit-identical to
// the level kernel and the CPU loop.
#include "common.cuh"
#include "kernels.cuh"
#include <algorithm>

namespace b200 {

constexpr long long SK_SPIN_LIMIT = 1LL << 24;

// ---- values: CRS order of the ILU factor -> the two per-step streams ---------------------------------------------------------------
__global__ void k_skew_fill(SkewGeom g, int n, const int *__restrict__ rows, const int *__restrict__ cols, const double *__restrict__ ilu,
                            double *__restrict__ SL, double *__restrict__ SU) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) sk_fill_row(g, i, rows, cols, ilu, SL, SU);
}

// guarded relaxed load: 0.0 when the operand does not exist (predicated, no branch)
__device__ __forceinline__ double ld_relaxed_if(const double *p, bool pred) {
  double v = 0.0;
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q ld.relaxed.gpu.global.f64 %0, [%1]; }" : "+d"(v) : "l"(p), "r"((unsigned)pred) : "memory");
  return v;
}

__global__ void k_skew_prepare(int n, double *a, double *b) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { a[i] = sent; b[i] = sent; }
}

// ---- the sweep ------------------------------------------------------------------------------------------------------------------------
//   forward  (UPPER = false): out_i = rhs_i - sum_{j<i} L_ij out_j                   (4642-4649)
//   backward (UPPER = true) : out_i = Dinv_i * (rhs_i - sum_{j>i} U_ij out_j)       (4653-4660)
template <bool UPPER>
__global__ void __launch_bounds__(128) k_skew(SkewGeom g, const double *__restrict__ S, const double *__restrict__ rhs, double *out,
                                              double *__restrict__ out2, Ctrl *ctrl) {
  if (ctrl->done) return;
  constexpr int NE = UPPER ? 14 : 13;
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long ntasks = g.ntasks();
  long long spins = 0;
  for (long long k = gw; k < ntasks; k += nw) {
    const int C = (int)(k / g.S), s = (int)(k % g.S);
    const int B0 = s * g.BW, nb = g.nb(s), nsteps = g.nsteps(s);
    const int Bq = B0 + lane;
    const double *Sp = S + g.step_base(C, s) * NE * 32 + lane;
    double h0 = 0.0, h1 = 0.0, h2 = 0.0;                             // own results of steps t-1, t-2, t-3
    double v[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) v[e] = ld_stream(Sp + (long long)e * 32);
    // previous-plane operands (slots 0..8) are requested ONE STEP AHEAD: values are written exactly once (sentinel -> result), so an
    // early read either is the result or still the sentinel, in which case the slot is polled again when it is needed
    double xn[9]; unsigned hasn = 0;
    {
      const int A0 = -2 * lane;
      const bool act0 = lane < nb && A0 >= 0 && A0 < g.NR;
#pragma unroll
      for (int e = 0; e < 9; ++e) {
        int dA, dB, dC; sk_offset(e, dA, dB, dC);
        const bool ex = act0 && g.inside(A0 + dA, Bq + dB, C + dC);
        xn[e] = ld_relaxed_if(out + (ex ? g.nat(UPPER, A0 + dA, Bq + dB, C + dC) : 0), ex);
        hasn |= (unsigned)ex << e;
      }
    }
    for (int t = 0; t < nsteps; ++t) {
      const int A = t - 2 * lane;
      const bool act = lane < nb && A >= 0 && A < g.NR;
      double vn[NE];                                                 // next step's entries in flight while this one computes
      if (t + 1 < nsteps) {
#pragma unroll
        for (int e = 0; e < NE; ++e) vn[e] = ld_stream(Sp + ((long long)(t + 1) * NE + e) * 32);
      } else {
#pragma unroll
        for (int e = 0; e < NE; ++e) vn[e] = 0.0;
      }
      // operands that live in L2: the previous plane (slots 0..8, requested last step) and, for lane 0, line B0-1 of this plane (9..11)
      double x[12]; unsigned has = hasn;
      const double *addr[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        int dA, dB, dC; sk_offset(e, dA, dB, dC);
        const bool ex = act && (e < 9 || lane == 0) && g.inside(A + dA, Bq + dB, C + dC);
        addr[e] = out + (ex ? g.nat(UPPER, A + dA, Bq + dB, C + dC) : 0);
        if (e >= 9) has |= (unsigned)ex << e;
      }
#pragma unroll
      for (int e = 0; e < 9; ++e) x[e] = xn[e];
#pragma unroll
      for (int e = 9; e < 12; ++e) x[e] = ld_relaxed_if(addr[e], (has >> e) & 1u);
      {                                                              // request the next step's previous-plane operands
        const int A1 = A + 1;
        const bool act1 = (t + 1 < nsteps) && lane < nb && A1 >= 0 && A1 < g.NR;
        hasn = 0;
#pragma unroll
        for (int e = 0; e < 9; ++e) {
          int dA, dB, dC; sk_offset(e, dA, dB, dC);
          const bool ex = act1 && g.inside(A1 + dA, Bq + dB, C + dC);
          xn[e] = ld_relaxed_if(out + (ex ? g.nat(UPPER, A1 + dA, Bq + dB, C + dC) : 0), ex);
          hasn |= (unsigned)ex << e;
        }
      }
      unsigned pend = 0;
#pragma unroll
      for (int e = 0; e < 12; ++e) pend |= (unsigned)(((has >> e) & 1u) && is_sentinel(x[e])) << e;
      int failed = 0;
      while (__any_sync(0xffffffffu, pend != 0)) {
        if (++spins > SK_SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        // a warp whose producers are far behind (pipeline fill) must not keep polling L2: back off up to ~4 us between rounds
        if (++failed > 2) __nanosleep(failed < 8 ? (32u << failed) : 4096u);
#pragma unroll
        for (int e = 0; e < 12; ++e)
          if ((pend >> e) & 1u) { x[e] = ld_relaxed(addr[e]); if (!is_sentinel(x[e])) pend &= ~(1u << e); }
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) if ((pend >> e) & 1u) x[e] = 0.0;   // only after a timeout: keep the arithmetic finite
      __syncwarp();
      // line B-1 of this plane from lane j-1: ages 3, 2, 1 are its rows A-1, A, A+1
      const double s3 = __shfl_up_sync(0xffffffffu, h2, 1), s2 = __shfl_up_sync(0xffffffffu, h1, 1), s1 = __shfl_up_sync(0xffffffffu, h0, 1);
      if (lane > 0) { x[9] = s3; x[10] = s2; x[11] = s1; }
      const long long i = act ? g.nat(UPPER, A, Bq, C) : 0;
      double acc = act ? rhs[i] : 0.0;
      if (!UPPER) {
#pragma unroll
        for (int e = 0; e < 12; ++e) acc = nfms(acc, v[e], x[e]);
        acc = nfms(acc, v[12], h0);
      } else {
        acc = nfms(acc, v[12], h0);
#pragma unroll
        for (int e = 11; e >= 0; --e) acc = nfms(acc, v[e], x[e]);
        acc = __dmul_rn(v[13], acc);
      }
      if (acc != acc) acc = __longlong_as_double((long long)CANON_NAN);
      h2 = h1; h1 = h0;
      if (act) {
        h0 = acc;
        st_relaxed(out + i, acc);
        if (out2) out2[i] = acc;
      }
#pragma unroll
      for (int e = 0; e < NE; ++e) v[e] = vn[e];
    }
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------------------------
void skew_release(Handle &h) {
  h.sk.SL.release(); h.sk.SU.release(); h.sk.y.release(); h.sk.x.release(); h.sk.ready = false; h.sk.tried = false;
}

void skew_analyse(Handle &h) {
  if (h.sk.ready || h.sk.tried) return;
  h.sk.tried = true;
  const char *why = nullptr;
  if (h.ilu_sep()) why = "ILU(n > 0) / BILU pattern";
  else if (h.nranks > 1) why = "partitioned handle";
  else why = sk_detect(h.n, h.h_rows.data(), h.h_cols.data(), h.h_diag.data(), h.sk.g);
  if (why) {
    if (getenv("B200_SKEW_DEBUG")) fprintf(stderr, "[skew] not usable (%s): level kernel stays\n", why);
    return;
  }
  const SkewGeom &g = h.sk.g;
  const size_t nl = (size_t)g.total_steps() * 13 * 32, nu = (size_t)g.total_steps() * 14 * 32;
  h.sk.SL.ensure(nl); h.sk.SU.ensure(nu);
  B200_CUDA(cudaMemsetAsync(h.sk.SL.p, 0, nl * sizeof(double), h.stream));
  B200_CUDA(cudaMemsetAsync(h.sk.SU.p, 0, nu * sizeof(double), h.stream));
  h.sk.y.ensure(std::max(h.n, 1)); h.sk.x.ensure(std::max(h.n, 1));
  h.sk.ready = true;
  if (getenv("B200_SKEW_DEBUG"))
    fprintf(stderr, "[skew] grid %d x %d x %d, %d strips of %d lines, %lld tasks, %lld steps, streams %.2f + %.2f GB\n", g.NR, g.NL, g.NP, g.S, g.BW,
            g.ntasks(), g.total_steps(), nl * 8e-9, nu * 8e-9);
}

void skew_refresh_values(Handle &h) {
  if (!h.sk.ready || h.n == 0) return;
  k_skew_fill<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, h.stream>>>(h.sk.g, h.n, h.d_rows.p, h.d_cols.p, h.d_ilu.p, h.sk.SL.p, h.sk.SU.p);
  B200_CUDA(cudaGetLastError());
}

template <bool UPPER>
static void skew_launch(Handle &h, const double *S, const double *rhs, double *out, double *out2) {
  const void *kern = (const void *)k_skew<UPPER>;
  int dev = 0, sms = 0, per_sm = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0));
  B200_REQUIRE(per_sm >= 1, "skewed-lane triangular solve: kernel does not fit on an SM");
  const int want_per_sm = std::max(1, std::min(per_sm, h.sk_blocks_per_sm > 0 ? h.sk_blocks_per_sm : 1));
  const long long ntasks = h.sk.g.ntasks();
  const int blocks = (int)std::max<long long>(1, std::min<long long>((long long)sms * want_per_sm, (ntasks + 3) / 4));
  SkewGeom g = h.sk.g; Ctrl *ctrl = h.ctrl.p;
  void *argv[] = {(void *)&g, (void *)&S, (void *)&rhs, (void *)&out, (void *)&out2, (void *)&ctrl};
  B200_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(128), argv, 0, h.stream));
}

void lu_apply_skew(Handle &h, double *u, const double *v) {
  B200_REQUIRE(h.sk.ready, "skewed-lane triangular solve without a plan");
  double *xo = (u == v) ? h.sk.x.p : u;
  double *x2 = (u == v) ? u : nullptr;
  k_skew_prepare<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, h.stream>>>(h.n, h.sk.y.p, xo);
  skew_launch<false>(h, h.sk.SL.p, v, h.sk.y.p, nullptr);
  skew_launch<true>(h, h.sk.SU.p, h.sk.y.p, xo, x2);
  B200_CUDA(cudaGetLastError());
  h.st_launch += 3; h.st_pcond++;
}

}  // namespace b200

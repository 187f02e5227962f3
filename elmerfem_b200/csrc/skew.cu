// Skewed-lane wavefront triangular solves for the ILU(0) factor of a structured-grid stencil -- CRS_LUSolve, fem/src/CRSMatrix.F90:4590-4663.
// Geometry, data layout and operand routing: skewgeom.h (shared with the CPU emulation tests/skew_harness.cpp).
//
// Why: the level kernel pays one L2 hand-off per dependency level, 2 x 1401 of them on the 200^3 heat problem.  Here one warp owns a
// STRIP of <= 29 consecutive lines of one grid plane; lane j solves row a = t - 2j of its line at step t, so the in-plane operands are
// the lane's own previous result and three results of lane j-1 (one shuffle per step).  The solve vectors live in the skewed
// (task, step, lane) layout of the matrix stream, so the operands of the previous plane are a coalesced row of the task below:
// ONE relaxed load per lane and step (requested E steps ahead, polled only when it still holds the sentinel) plus two shuffles; the
// 13 (14) matrix entries and the right-hand side of a step arrive by TMA bulk copies into a per-warp shared-memory ring.
// Tasks (plane, strip) are taken in increasing order by a co-resident grid and every dependency points to a lower task: no deadlock
// for any grid shape.  Arithmetic: the reference's operations in the reference's order (entries in ascending column order, separate
// multiply / subtract roundings, inverse diagonal last); pad entries (neighbours outside the grid) are 0 x 0.  Bit-identical to the
// level kernel and to the CPU loop.
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

// natural order -> skewed layout (right-hand side of the forward sweep), and the sentinel fill of the two result vectors
__global__ void k_skew_in(SkewGeom g, int n, const double *__restrict__ v, double *__restrict__ yin, long long nv, double *__restrict__ a,
                          double *__restrict__ b) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = t0; i < nv; i += stride) { a[i] = sent; b[i] = sent; }
  for (long long i = t0; i < n; i += stride) yin[g.vslot(i)] = v[i];
}
// skewed layout -> natural order
__global__ void k_skew_out(SkewGeom g, int n, const double *__restrict__ x, double *__restrict__ u) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) u[i] = x[g.vslot(i)];
}

// guarded relaxed load: 0.0 when the operand does not exist (predicated, no branch)
__device__ __forceinline__ double ld_relaxed_if(const double *p, bool pred) {
  double v = 0.0;
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q ld.relaxed.gpu.global.f64 %0, [%1]; }" : "+d"(v) : "l"(p), "r"((unsigned)pred));
  return v;
}
// the sweep's own store: ordered with the other volatile accesses of the sweep, but no compiler barrier for the shared-memory reads
__device__ __forceinline__ void st_relaxed_nc(double *p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v)); }
__device__ __forceinline__ unsigned sk_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sk_mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void sk_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sk_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool sk_mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ long long sk_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// ---- the sweep ------------------------------------------------------------------------------------------------------------------------
//   forward  (UPPER = false): out_i = rhs_i - sum_{j<i} L_ij out_j                   (4642-4649)
//   backward (UPPER = true) : out_i = Dinv_i * (rhs_i - sum_{j>i} U_ij out_j)       (4653-4660)
// S: matrix stream, RHS / Y: right-hand side and result in the skewed layout (Y pre-filled with the sentinel).
// E: steps a result-vector load is requested ahead; HN = E + 5 values of history per lane (ages 0 .. E+4); G: steps per unrolled group
// (a multiple of HN, so that history positions are compile-time constants and no register with a load in flight is ever copied);
// TG: steps per TMA chunk, NSLOT chunks in flight.
template <bool UPPER, int E, int G, int TG, int NSLOT>
__global__ void __launch_bounds__(128) k_skew(SkewGeom g, const double *__restrict__ S, const double *__restrict__ RHS, double *Y, Ctrl *ctrl,
                                              long long *trace) {
  if (ctrl->done) return;
  constexpr int NE = UPPER ? 14 : 13, HN = E + 5;
  constexpr int STEP_D = NE * 32, CH_M = TG * STEP_D, CH_D = CH_M + TG * 32;     // doubles per chunk: matrix part, + rhs part
  constexpr unsigned FULL = 0xffffffffu;
  static_assert(G % HN == 0 && G % TG == 0, "group size must be a multiple of the history length and of the TMA chunk");
  extern __shared__ __align__(128) unsigned char sk_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double *ring = reinterpret_cast<double *>(sk_smem) + (size_t)warp * NSLOT * CH_D;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(sk_smem + (size_t)wpb * NSLOT * CH_D * sizeof(double)) + warp * NSLOT;
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) sk_mbar_init(sk_smem_u32(bars + q), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const long long gw = (long long)blockIdx.x * wpb + warp, nw = (long long)gridDim.x * wpb;
  const long long ntasks = g.ntasks();
  const int NR = g.NR;
  long long spins = 0;
  unsigned nch = 0;                                                   // chunks consumed so far by this warp (ring position / parity)
  for (long long k = gw; k < ntasks; k += nw) {
    const SkewTask T = sk_task(g, UPPER, k);
    const int nb = T.nb, nsteps = T.nsteps;
    const int ngroups = (nsteps + G - 1) / G, nchunks = ngroups * (G / TG);
    long long tr_start = 0, tr_slow = 0, tr_polls = 0, tr_cyc_tma = 0, tr_cyc_poll = 0;
    if (trace) tr_start = sk_gtime();
    auto issue = [&](int ci, unsigned slot) {                         // chunk ci of this task -> ring slot (lane 0 only)
      const long long first = sk_group_first(T, UPPER, ci * TG, TG);
      const unsigned bar = sk_smem_u32(bars + slot);
      sk_mbar_expect_tx(bar, CH_D * 8);
      sk_bulk_g2s(sk_smem_u32(ring + (size_t)slot * CH_D), S + first * STEP_D, CH_M * 8, bar);
      sk_bulk_g2s(sk_smem_u32(ring + (size_t)slot * CH_D + CH_M), RHS + first * 32, TG * 32 * 8, bar);
    };
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) if (q < nchunks) issue(q, (nch + q) % NSLOT);
    }
    // the lane's source in the result vector, its own output row, its partners in the three shuffles
    const SkewSrc src = sk_source(g, UPPER, k, lane);
    constexpr int vstride = UPPER ? -32 : 32;
    const double *psrc = Y + src.idx0;
    double *pown = Y + sk_vidx(T, UPPER, 0, lane < nb ? lane : 0);
    const bool lane_on = lane < nb;
    const int src_near = lane < 31 ? lane + 1 : 31, src_far = lane == 0 ? 31 : lane - 1, src_h = lane == 0 ? 30 : lane - 1;
    const int lpos = UPPER ? (lane_on ? nb - 1 - lane : lane) : lane;  // the lane's position inside a 32-wide row of the ring
    const int rlim = NR + 2 * lane;                                    // rows of this lane: steps 2 lane .. rlim - 1
    double H[HN];
#pragma unroll
    for (int q = 0; q < HN; ++q) H[q] = 0.0;
    double h0 = 0.0, wm = 0.0, w0 = 0.0, um = 0.0, u0 = 0.0, up = 0.0, dm = 0.0, d0 = 0.0, dp = 0.0;
    // what the load stage of a step hands to its tail one iteration later
    double part = 0.0, k11 = 0.0, k12 = 0.0, k13 = 0.0, P[11];
#pragma unroll
    for (int q = 0; q < 11; ++q) P[q] = 0.0;
    // Operand windows of step t (u = position of t in its unrolled group, t == u modulo HN): request the value of step t + E, take
    // the one of step t (polling where the sentinel is still there), exchange it with the neighbour lanes.
    auto advance = [&](int t, int u) {
      {
        const int tau = t + E + src.off;
        H[u % HN] = ld_relaxed_if(psrc + (long long)tau * vstride, (unsigned)(tau - src.lo) < (unsigned)src.len);
      }
      double head = H[(u + HN - E) % HN];                             // requested E steps ago, needed now
      if (__any_sync(FULL, is_sentinel(head))) {
        ++tr_slow;
        const long long c0 = trace ? clock64() : 0;
        const double *pa = psrc + (long long)(t + src.off) * vstride;
        bool bad = is_sentinel(head);
        do {
          if (++spins > SK_SPIN_LIMIT) { ctrl->spin_timeout = 1; bad = false; head = 0.0; }
          ++tr_polls;
          if (bad) { head = ld_relaxed(pa); bad = is_sentinel(head); }
        } while (__any_sync(FULL, bad));
        H[(u + HN - E) % HN] = head;
        if (trace) tr_cyc_poll += clock64() - c0;
      }
      if (lane == 30) h0 = head;                                      // lane 0's in-plane operand (a+1, b-1) of this step
      const double nearv = __shfl_sync(FULL, head, src_near);
      const double farv = __shfl_sync(FULL, H[(u + HN - E - 4) % HN], src_far);
      um = u0; u0 = up; up = farv;
      dm = d0; d0 = dp; dp = nearv;
    };
    // Tail of step t: the two operands that depend on step t-1 (own result, lane j-1's result), then the store.  The reference's
    // left-to-right subtraction puts them LAST in the forward sweep (3 dependent operations per step) and FIRST in the backward sweep.
    // (a+1, b-1) of step t = lane j-1's result of step t-1 (lane 30's fetch for lane 0): the one shuffle on the step-to-step recurrence
    auto neighbour = [&](int t) {
      double wn = __shfl_sync(FULL, h0, src_h);
      if (t + 1 >= rlim) wn = 0.0;                                    // neighbour outside the grid: pad entry times a clean zero
      return wn;
    };
    auto tail = [&](int t, double wn) {
      double acc;
      if (!UPPER) {
        acc = nfms(part, k11, wn);
        acc = nfms(acc, k12, h0);
      } else {
        acc = nfms(part, k12, h0);
        acc = nfms(acc, k11, wn);
#pragma unroll
        for (int e = 10; e >= 0; --e) acc = __dsub_rn(acc, P[e]);
        acc = __dmul_rn(k13, acc);
      }
      if (acc != acc) acc = __longlong_as_double((long long)CANON_NAN);
      if (lane_on && t >= 2 * lane && t < rlim) {
        h0 = acc;
        st_relaxed_nc(pown + (long long)t * vstride, acc);
      }
    };
#pragma unroll
    for (int u = 0; u < G; ++u) advance(u - G, u);                    // fills the windows; every lane is still before its first row
    for (int gi = 0; gi < ngroups; ++gi) {
#pragma unroll
      for (int u = 0; u < G; ++u) {
        const int t = gi * G + u;                                     // steps past nsteps (group padding) find every lane inactive
        if (u % TG == 0) {
          const unsigned slot = nch % NSLOT, bar = sk_smem_u32(bars + slot), parity = (nch / NSLOT) & 1u;
          const long long c0 = trace ? clock64() : 0;
          while (!sk_mbar_try_wait(bar, parity)) { if (++spins > SK_SPIN_LIMIT) { ctrl->spin_timeout = 1; break; } }
          if (trace) tr_cyc_tma += clock64() - c0;
        }
        const double wn = neighbour(t - 1);                            // before advance() refills lane 30
        advance(t, u);
        // From here to the end of the iteration is one basic block: the tail of step t-1 (the recurrence) and the load stage of step t
        // (everything that does not depend on step t-1) are independent instruction streams the scheduler interleaves.
        tail(t - 1, wn);
        wm = w0; w0 = wn;                                             // in-plane operands (a-1, b-1), (a, b-1) of step t
        const double *cp = ring + (size_t)(nch % NSLOT) * CH_D;
        const int spos = UPPER ? TG - 1 - (u % TG) : (u % TG);
        const double *vp = cp + spos * STEP_D + lpos;
        const double rv = cp[CH_M + spos * 32 + lpos];
        double v[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) v[e] = vp[e * 32];
        const double om = H[(u + HN - E - 4) % HN], o0 = H[(u + HN - E - 3) % HN], op = H[(u + HN - E - 2) % HN];
        const double x[11] = {um, u0, up, om, o0, op, dm, d0, dp, wm, w0};
        k11 = v[11]; k12 = v[12];
        if (!UPPER) {
          part = rv;
#pragma unroll
          for (int e = 0; e < 11; ++e) part = nfms(part, v[e], x[e]);
        } else {
          part = rv; k13 = v[13];
#pragma unroll
          for (int e = 0; e < 11; ++e) P[e] = __dmul_rn(v[e], x[e]);
        }
        if (u % TG == TG - 1) {
          __syncwarp();
          const int ci = gi * (G / TG) + u / TG;                      // chunk just consumed
          if (lane == 0 && ci + NSLOT < nchunks) issue(ci + NSLOT, nch % NSLOT);
          ++nch;
        }
      }
    }
    tail(ngroups * G - 1, neighbour(ngroups * G - 1));
    if (trace && lane == 0) {
      unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      long long *r = trace + ((UPPER ? ntasks : 0) + k) * 8;
      r[0] = tr_start; r[1] = sk_gtime(); r[2] = tr_slow; r[3] = tr_polls; r[4] = smid; r[5] = tr_cyc_tma; r[6] = tr_cyc_poll;
    }
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------------------------
constexpr int SK_PAD_STEPS = 16;      // layout padding in front of and behind the streams: the first / last TMA chunk of a task may overhang

void skew_release(Handle &h) {
  h.sk.SL.release(); h.sk.SU.release(); h.sk.yin.release(); h.sk.y.release(); h.sk.x.release(); h.sk.trace.release(); h.sk.ready = false; h.sk.tried = false;
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
  const size_t steps = (size_t)g.total_steps() + 2 * SK_PAD_STEPS;
  h.sk.SL.ensure(steps * 13 * 32); h.sk.SU.ensure(steps * 14 * 32);
  B200_CUDA(cudaMemsetAsync(h.sk.SL.p, 0, steps * 13 * 32 * sizeof(double), h.stream));
  B200_CUDA(cudaMemsetAsync(h.sk.SU.p, 0, steps * 14 * 32 * sizeof(double), h.stream));
  h.sk.yin.ensure(steps * 32); h.sk.y.ensure(steps * 32); h.sk.x.ensure(steps * 32);
  B200_CUDA(cudaMemsetAsync(h.sk.yin.p, 0, steps * 32 * sizeof(double), h.stream));
  B200_CUDA(cudaMemsetAsync(h.sk.y.p, 0, steps * 32 * sizeof(double), h.stream));
  B200_CUDA(cudaMemsetAsync(h.sk.x.p, 0, steps * 32 * sizeof(double), h.stream));
  h.sk.ready = true;
  if (getenv("B200_SKEW_DEBUG"))
    fprintf(stderr, "[skew] grid %d x %d x %d, %d strips of %d lines, %lld tasks, %lld steps, streams %.2f + %.2f GB\n", g.NR, g.NL, g.NP, g.S, g.BW,
            g.ntasks(), g.total_steps(), steps * 13 * 32 * 8e-9, steps * 14 * 32 * 8e-9);
}

void skew_refresh_values(Handle &h) {
  if (!h.sk.ready || h.n == 0) return;
  k_skew_fill<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, h.stream>>>(h.sk.g, h.n, h.d_rows.p, h.d_cols.p, h.d_ilu.p,
                                                                             h.sk.SL.p + (size_t)SK_PAD_STEPS * 13 * 32, h.sk.SU.p + (size_t)SK_PAD_STEPS * 14 * 32);
  B200_CUDA(cudaGetLastError());
}

template <bool UPPER, int E, int G, int TG, int NSLOT>
static void skew_launch_cfg(Handle &h, int wpb, const double *S, const double *rhs, double *out) {
  const void *kern = (const void *)k_skew<UPPER, E, G, TG, NSLOT>;
  constexpr int NE = UPPER ? 14 : 13;
  const size_t per_warp = (size_t)NSLOT * TG * (NE + 1) * 32 * sizeof(double) + NSLOT * sizeof(unsigned long long);
  wpb = (int)std::max<size_t>(1, std::min<size_t>(wpb, (227 * 1024 - 1024) / per_warp));
  const size_t smem = per_warp * wpb;
  int dev = 0, sms = 0, per_sm = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem));
  B200_REQUIRE(per_sm >= 1, "skewed-lane triangular solve: kernel does not fit on an SM");
  const int want_per_sm = h.sk_blocks_per_sm > 0 ? std::min(per_sm, h.sk_blocks_per_sm) : per_sm;
  const long long ntasks = h.sk.g.ntasks();
  const int blocks = (int)std::max<long long>(1, std::min<long long>((long long)sms * want_per_sm, (ntasks + wpb - 1) / wpb));
  SkewGeom g = h.sk.g; Ctrl *ctrl = h.ctrl.p; long long *trace = h.sk.trace_on ? h.sk.trace.p : nullptr;
  void *argv[] = {(void *)&g, (void *)&S, (void *)&rhs, (void *)&out, (void *)&ctrl, (void *)&trace};
  B200_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(wpb * 32), argv, smem, h.stream));
}

template <bool UPPER>
static void skew_launch(Handle &h, const double *S, const double *rhs, double *out) {
  const int wpb = std::max(1, std::min(4, h.sk_wpb > 0 ? h.sk_wpb : 4));
  switch (h.sk_cfg) {
    case 1: skew_launch_cfg<UPPER, 3, 8, 8, 2>(h, wpb, S, rhs, out); break;
    case 2: skew_launch_cfg<UPPER, 1, 6, 3, 4>(h, wpb, S, rhs, out); break;
    case 3: skew_launch_cfg<UPPER, 7, 12, 4, 4>(h, wpb, S, rhs, out); break;
    case 4: skew_launch_cfg<UPPER, 3, 8, 4, 3>(h, wpb, S, rhs, out); break;
    default: skew_launch_cfg<UPPER, 3, 8, 4, 4>(h, wpb, S, rhs, out); break;
  }
}

void lu_apply_skew(Handle &h, double *u, const double *v) {
  B200_REQUIRE(h.sk.ready, "skewed-lane triangular solve without a plan");
  const SkewGeom &g = h.sk.g;
  const size_t pad = (size_t)SK_PAD_STEPS * 32;
  double *yin = h.sk.yin.p + pad, *y = h.sk.y.p + pad, *x = h.sk.x.p + pad;
  const int blocks = std::min((h.n + 255) / 256, NUM_SMS * 8);
  k_skew_in<<<blocks, 256, 0, h.stream>>>(g, h.n, v, yin, g.vlen(), y, x);
  skew_launch<false>(h, h.sk.SL.p + (size_t)SK_PAD_STEPS * 13 * 32, yin, y);
  skew_launch<true>(h, h.sk.SU.p + (size_t)SK_PAD_STEPS * 14 * 32, y, x);
  k_skew_out<<<blocks, 256, 0, h.stream>>>(g, h.n, x, u);
  B200_CUDA(cudaGetLastError());
  h.st_launch += 4; h.st_pcond++;
}

// per-task trace (profiles/tools/skew_lab.cu): 8 long long per (sweep, task)
void skew_trace_enable(Handle &h, bool on) {
  if (on) {
    const size_t m = (size_t)h.sk.g.ntasks() * 2 * 8;
    h.sk.trace.ensure(m);
    B200_CUDA(cudaMemsetAsync(h.sk.trace.p, 0, m * sizeof(long long), h.stream));
  }
  h.sk.trace_on = on;
}
void skew_trace_fetch(Handle &h, std::vector<long long> &out) {
  out.assign((size_t)h.sk.g.ntasks() * 2 * 8, 0);
  B200_CUDA(cudaStreamSynchronize(h.stream));
  B200_CUDA(cudaMemcpy(out.data(), h.sk.trace.p, out.size() * sizeof(long long), cudaMemcpyDeviceToHost));
}

}  // namespace b200

// Lane-tile triangular solves for the ILU(0) factor of a structured-grid stencil -- CRS_LUSolve, fem/src/CRSMatrix.F90:4590-4663.
// Geometry, data layout, shuffle routing and the row program: lanegeom.h (shared with the CPU emulation tests/lane_harness.cpp).
//
// Why: the dependency DAG of the 27-point factor has NR + 2 NL + 4 NP levels (1401 on the 200^3 heat problem) and every level-to-level
// hand-off that goes through L2 (level kernel: 0.9 us) or through shared memory + a CTA barrier (wave tiles: 0.3-0.55 us) sits on the
// critical path of the sweep.  Here a tile is ONE WARP: lane = sheared grid line, TC planes per lane, lanes skewed by two steps, so that
// every operand of a row is a register of the lane or arrives by one warp shuffle from lane j-1 (lanegeom.h).  A step (= one dependency
// level of the tile) is a few shuffles, TC x 13 multiply-subtract pairs in the reference's order and TC coalesced stores: no barrier, no
// shared-memory hand-off.  Only the lines just outside a tile (two ghost lanes, one plane below) come from L2, requested LT_E steps ahead
// with the sentinel protocol; neighbouring tiles run concurrently, a few steps apart.
// Matrix entries + right-hand sides of a tile step are one contiguous block each; lane 0 of the warp moves them with two bulk copies
// (cp.async.bulk, mbarrier completion) into a private D-slot shared-memory ring D steps ahead, lane 1 prefetches the stream into L2
// further ahead.  Arithmetic: the reference's operations in the reference's order, separate roundings; pad entries are (+0) x (+0).
// Bit-identical to the level kernel and to the CPU loop.
#include "common.cuh"
#include "kernels.cuh"
#include "lanegeom.h"
#include <algorithm>
#include <type_traits>

namespace b200 {

constexpr long long LT_SPIN_LIMIT = 1LL << 23;
constexpr int LT_PF = 24;       // steps the L2 prefetch runs ahead

__global__ void k_lane_fill(LaneGeom g, const int *__restrict__ tile_of, int n, const int *__restrict__ rows, const int *__restrict__ cols,
                            const double *__restrict__ ilu, double *__restrict__ SL, double *__restrict__ SU) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) lt_fill_row(g, tile_of, i, rows, cols, ilu, SL, SU);
}
// natural order -> right-hand-side rows of the forward stream + sentinel fill of the forward result
__global__ void k_lane_in(LaneGeom g, const int *__restrict__ tile_of, int n, const double *__restrict__ v, double *__restrict__ SL, long long nv,
                          double *__restrict__ y) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = t0; i < nv; i += stride) y[i] = sent;
  for (long long i = t0; i < n; i += stride) {
    const int a = (int)(i % g.NR), b = (int)((i / g.NR) % g.NL), c = (int)(i / ((long long)g.NR * g.NL));
    SL[lt_rhs_index(lt_pos(g, tile_of, a, b, c), LT_ROWS_L)] = v[i];
  }
}
// tile layout -> natural order; the slots are handed back as sentinels for the next application
__global__ void k_lane_out(LaneGeom g, const int *__restrict__ tile_of, int n, double *__restrict__ x, double *__restrict__ u) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(i % g.NR), b = (int)((i / g.NR) % g.NL), c = (int)(i / ((long long)g.NR * g.NL));
    const long long p = lt_pos(g, tile_of, a, b, c);
    u[i] = x[p]; x[p] = sent;
  }
}
__global__ void k_lane_sentinel(long long nv, double *__restrict__ x) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) x[i] = sent;
}

__device__ __forceinline__ unsigned lt_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lt_mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void lt_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lt_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void lt_mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LT_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra LT_DONE;\n"
      "bra LT_WAIT;\n"
      "LT_DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool lt_mbar_try(unsigned bar, unsigned parity) {     // may suspend for a bounded time
  unsigned ok;
  asm volatile("{\n .reg .pred P1;\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n selp.u32 %0, 1, 0, P1;\n }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool lt_mbar_test(unsigned bar, unsigned parity) {    // never suspends
  unsigned ok;
  asm volatile("{\n .reg .pred P1;\n mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n selp.u32 %0, 1, 0, P1;\n }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void lt_mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void lt_prefetch_l2(const void *p, unsigned bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory"); }
__device__ __forceinline__ void lt_st_relaxed(double *p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v)); }
__device__ __forceinline__ double lt_ld_relaxed(const double *p) {
  double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v;
}
// predicated forms: no branch around the access
__device__ __forceinline__ double lt_ld_relaxed_if(const double *p, bool ok) {
  double v;
  asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n mov.f64 %0, 0d0000000000000000;\n @q ld.relaxed.gpu.global.f64 %0, [%1];\n }" : "=d"(v) : "l"(p), "r"((int)ok) : "memory");
  return v;
}
__device__ __forceinline__ void lt_st_relaxed_if(double *p, double v, bool ok) {
  asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q st.relaxed.gpu.global.f64 [%0], %1;\n }" ::"l"(p), "d"(v), "r"((int)ok));
}
__device__ __forceinline__ void lt_st_if(double *p, double v, bool ok) {
  asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q st.global.f64 [%0], %1;\n }" ::"l"(p), "d"(v), "r"((int)ok));
}
__device__ __forceinline__ long long lt_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// ---- the sweep ------------------------------------------------------------------------------------------------------------------------
//   forward  (UPPER = false): out_i = rhs_i - sum_{j<i} L_ij out_j                   (4642-4649)
//   backward (UPPER = true) : out_i = Dinv_i * (rhs_i - sum_{j>i} U_ij out_j)       (4653-4660)
// S: this sweep's stream (entries + right-hand sides at pos(sweep coordinates)); Q: result at pos(mirrored sweep coordinates), pre-filled
// with the sentinel; R2: the backward stream, whose right-hand-side rows the forward sweep fills (nullptr for the backward sweep).
// D: ring depth (slots per warp).  Block = W consumer warps, one tile each at a time, + one producer warp whose lane w feeds the ring of
// consumer w with one bulk copy per tile step (full / empty mbarrier pair per slot).  Consumer (block, w) takes tiles block + grid * w,
// + grid * W, ...: consecutive tiles (neighbouring start levels) sit on different SMs.
template <bool UPPER, int TC>
__global__ void __launch_bounds__(256, 1) k_lane(LaneGeom g, const int *__restrict__ tile_of, const int *__restrict__ tile_sig, const int *__restrict__ tile_grp,
                                                 const double *__restrict__ S, double *Q, double *R2, Ctrl *ctrl, int D, long long *trace) {
  if (ctrl->done) return;
  constexpr int NE = UPPER ? 14 : 13, NROW = NE + 1;
  constexpr unsigned SLOT = TC * NROW * 256u;                       // bytes of a tile step
  constexpr long long BLKD = TC * NROW * 32;                        // doubles of a tile step
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char lt_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, W = (blockDim.x >> 5) - 1;
  const unsigned ring0 = lt_smem_u32(lt_smem), full0 = ring0 + (unsigned)(W * D) * SLOT, empty0 = full0 + (unsigned)(W * D) * 8u;
  if (threadIdx.x == 0) {
    for (int d = 0; d < W * D; ++d) { lt_mbar_init(full0 + 8u * d, 1); lt_mbar_init(empty0 + 8u * d, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int NT = g.NT, NR = g.NR, ntiles = g.ntiles, kstep = gridDim.x * W;

  if (wib == W) {
    // ---------------- producer warp: lane w serves consumer w ----------------
    bool act = lane < W;
    int k = blockIdx.x + gridDim.x * lane;
    if (k >= ntiles) act = false;
    int tau = 0, slot = 0; unsigned par = 0;
    const unsigned ring = ring0 + (unsigned)(lane * D) * SLOT, full = full0 + (unsigned)(lane * D) * 8u, empty = empty0 + (unsigned)(lane * D) * 8u;
    while (__any_sync(FULL, act)) {
      if (act && lt_mbar_test(empty + 8u * slot, par ^ 1u)) {
        const double *src = S + ((long long)k * NT + tau) * BLKD;
        lt_mbar_expect_tx(full + 8u * slot, SLOT);
        lt_bulk_g2s(ring + (unsigned)slot * SLOT, src, SLOT, full + 8u * slot);
        if ((tau & 3) == 0) {                                       // L2 prefetch LT_PF steps ahead, into the next tile of this consumer at the end
          int tp = tau + LT_PF, kp = k;
          if (tp >= NT) { tp -= NT; kp += kstep; }
          if (kp < ntiles && tp + 4 <= NT) lt_prefetch_l2(S + ((long long)kp * NT + tp) * BLKD, 4u * SLOT);
        }
        if (++slot == D) { slot = 0; par ^= 1u; }
        if (++tau == NT) { tau = 0; k += kstep; if (k >= ntiles) act = false; }
      }
    }
    return;
  }

  // ---------------- consumer warp ----------------
  const unsigned ring = ring0 + (unsigned)(wib * D) * SLOT, full = full0 + (unsigned)(wib * D) * 8u, empty = empty0 + (unsigned)(wib * D) * 8u;
  const unsigned char *ring_g = lt_smem + (size_t)(wib * D) * SLOT;
  (void)ring;
  const long long stride = g.stride();
  int slot = 0; unsigned par = 0;                                   // ring position / phase of the next step to load
  long long spins = 0;
  for (int k = blockIdx.x + gridDim.x * wib; k < ntiles; k += kstep) {
    const int sig = tile_sig[k], C = tile_grp[k];
    // Per plane slot: pointer to the line's row of the CURRENT step in the result vector (row a = tau - 2 lane - 2 p; one subtraction
    // per step), and the row range in which the lane replays (nrq) or solves and stores (nrs) it -- 0 where it does neither.
    double *qp[TC + 1], *rp[TC + 1]; unsigned nrq[TC + 1], nrs[TC + 1];
#pragma unroll
    for (int p = 0; p <= TC; ++p) {
      const LaneLine ln = lt_line(g, sig, C, lane, p);
      const bool rep = lt_replayed(lane, p);
      nrq[p] = (ln.valid && rep) ? (unsigned)NR : 0u;
      nrs[p] = (ln.valid && !rep) ? (unsigned)NR : 0u;
      const long long q0 = ln.valid ? lt_pos_mirror(g, tile_of, 0, ln.b, ln.c) : 0;
      qp[p] = Q + q0 + (long long)(2 * lane + 2 * p) * stride;
      rp[p] = UPPER ? nullptr : R2 + lt_rhs_index(q0, LT_ROWS_U) + (long long)(2 * lane + 2 * p) * (TC * LT_ROWS_U * 32);
    }
    const long long eoff = (long long)LT_E * stride;
    int ab = -2 * lane;                                             // tau - 2 lane
    long long tr0 = 0, tr_polls = 0, tr_mbar = 0, tr_poll = 0, tr_first = 0, ph[5] = {0, 0, 0, 0, 0}, pc = 0;
    if (trace && lane == 0) tr0 = lt_gtime();
    LaneHist<TC> h;
    lt_hist_clear(h);
    double Hh[TC + 1][LT_E + 1];
#pragma unroll
    for (int p = 0; p <= TC; ++p) {
#pragma unroll
      for (int u = 0; u < LT_E; ++u) Hh[p][u] = lt_ld_relaxed_if(qp[p] - (long long)u * stride, (unsigned)(ab + u - 2 * p) < nrq[p]);
      Hh[p][LT_E] = 0.0;
    }
    double V[TC][NROW];                                             // entries + right-hand side of the step about to run
    auto load_step = [&](bool ready) {                              // ring slot -> registers, slot handed back to the producer
      if (!ready) {
        if (trace) { const long long c0 = clock64(); lt_mbar_wait(full + 8u * slot, par); tr_mbar += clock64() - c0; }
        else lt_mbar_wait(full + 8u * slot, par);
      }
      const double *sl = reinterpret_cast<const double *>(ring_g + (size_t)slot * SLOT) + lane;
#pragma unroll
      for (int p = 0; p < TC; ++p) {
#pragma unroll
        for (int e = 0; e < NROW; ++e) V[p][e] = sl[(p * NROW + e) * 32];
      }
      __syncwarp();
      if (lane == 0) lt_mbar_arrive(empty + 8u * slot);
      if (++slot == D) { slot = 0; par ^= 1u; }
    };
    load_step(false);
    auto step = [&](auto Uc, int tau) {
      constexpr int U = decltype(Uc)::value;
      if (trace) pc = clock64();
      // the two shuffles
      LaneMsg<TC> m;
      lt_send<TC, U>(h, m);
#pragma unroll
      for (int p = 0; p <= TC; ++p) m.r[p] = __shfl_up_sync(FULL, m.r[p], 1);
#pragma unroll
      for (int p = 0; p < TC; ++p) m.t[p] = __shfl_up_sync(FULL, m.t[p], 1);
      lt_recv<TC, U>(h, m);
      if (trace) { const long long c = clock64(); ph[0] += c - pc; pc = c; }
      // replayed values of step tau + LT_E; is the next step's block there?
#pragma unroll
      for (int p = 0; p <= TC; ++p) Hh[p][(U + LT_E) & LT_E] = lt_ld_relaxed_if(qp[p] - eoff, (unsigned)(ab + LT_E - 2 * p) < nrq[p]);
      const bool more = tau + 1 < NT;
      const bool ready = more && lt_mbar_try(full + 8u * slot, par);
      if (trace) { const long long c = clock64(); ph[1] += c - pc; pc = c; }
      // rows
      double out[TC + 1];
      out[0] = 0.0;
#pragma unroll
      for (int p = 1; p <= TC; ++p) {
        double acc = lt_row<UPPER, TC, U>(h, p, V[p - 1], V[p - 1][NE]);
        if (acc != acc) acc = __longlong_as_double((long long)CANON_NAN);
        const bool active = (unsigned)(ab - 2 * p) < nrs[p];
        if (!active) acc = 0.0;
        out[p] = acc;
        lt_st_relaxed_if(qp[p], acc, active);
        if (!UPPER) lt_st_if(rp[p], acc, active);
      }
      if (trace) { const long long c = clock64(); ph[2] += c - pc; pc = c; }
      if (more) load_step(ready);
      if (trace) { const long long c = clock64(); ph[3] += c - pc; pc = c; }
      // replayed values of this step (requested LT_E steps ago; a producer that is not that far ahead yet is polled)
      double hv[TC + 1];
      bool need = false;
#pragma unroll
      for (int p = 0; p <= TC; ++p) { hv[p] = Hh[p][U & LT_E]; need = need || is_sentinel(hv[p]); }
      if (__any_sync(FULL, need)) {
        const long long c0 = trace ? clock64() : 0;
        unsigned tries = 0;
        do {
          need = false;
#pragma unroll
          for (int p = 0; p <= TC; ++p)
            if (is_sentinel(hv[p])) { hv[p] = lt_ld_relaxed(qp[p]); need = need || is_sentinel(hv[p]); ++tr_polls; }
          if (need) {
            if (++spins > LT_SPIN_LIMIT) {
              ctrl->spin_timeout = 1;
#pragma unroll
              for (int p = 0; p <= TC; ++p) if (is_sentinel(hv[p])) hv[p] = 0.0;
              need = false;
            } else if (++tries > 8) __nanosleep(tries > 64 ? 400 : 100);
          }
        } while (__any_sync(FULL, need));
        if (trace) { tr_poll += clock64() - c0; if (tau == 0) tr_first = lt_gtime(); }
      }
#pragma unroll
      for (int p = 0; p <= TC; ++p) {
        h.X[p][U & 7] = (p == 0 || lane < LT_GH) ? hv[p] : out[p];
        qp[p] -= stride;
        if (!UPPER) rp[p] -= TC * LT_ROWS_U * 32;
      }
      ++ab;
      if (trace) { const long long c = clock64(); ph[4] += c - pc; pc = c; }
    };
    for (int t0 = 0; t0 < NT; t0 += 8) {
      step(std::integral_constant<int, 0>{}, t0); step(std::integral_constant<int, 1>{}, t0 + 1);
      step(std::integral_constant<int, 2>{}, t0 + 2); step(std::integral_constant<int, 3>{}, t0 + 3);
      step(std::integral_constant<int, 4>{}, t0 + 4); step(std::integral_constant<int, 5>{}, t0 + 5);
      step(std::integral_constant<int, 6>{}, t0 + 6); step(std::integral_constant<int, 7>{}, t0 + 7);
    }
    if (trace) {
      long long *r = trace + ((UPPER ? g.ntiles : 0) + (long long)k) * 16;
      if (lane == 0) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); r[0] = tr0; r[1] = lt_gtime(); r[3] = smid; r[4] = tr_mbar; r[5] = tr_poll; r[6] = tr_first; for (int q = 0; q < 5; ++q) r[8 + q] = ph[q]; }
      if (tr_polls) atomicAdd((unsigned long long *)(r + 2), (unsigned long long)tr_polls);
    }
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------------------------
void lane_release(Handle &h) {
  LanePlan &w = h.lt;
  w.SL.release(); w.SU.release(); w.y.release(); w.x.release(); w.tile_of.release(); w.tile_sig.release(); w.tile_grp.release(); w.trace.release();
  w.ready = false; w.tried = false;
}

void lane_analyse(Handle &h) {
  LanePlan &w = h.lt;
  if (w.ready || w.tried) return;
  w.tried = true;
  const char *why = nullptr;
  SkewGeom sg;
  if (h.ilu_sep()) why = "ILU(n > 0) / BILU pattern";
  else if (h.nranks > 1) why = "partitioned handle";
  else why = sk_detect(h.n, h.h_rows.data(), h.h_cols.data(), h.h_diag.data(), sg);
  if (why) {
    if (getenv("B200_WAVE_DEBUG")) fprintf(stderr, "[lane] not usable (%s): level kernel stays\n", why);
    return;
  }
  const int TC = std::max(1, std::min(4, h.lt_tc));
  LaneTiles T;
  lt_plan(w.g, sg.NR, sg.NL, sg.NP, TC, T);
  const LaneGeom &g = w.g;
  w.tile_of.ensure(T.tile_of.size()); w.tile_sig.ensure(T.sig.size()); w.tile_grp.ensure(T.grp.size());
  B200_CUDA(cudaMemcpyAsync(w.tile_of.p, T.tile_of.data(), T.tile_of.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaMemcpyAsync(w.tile_sig.p, T.sig.data(), T.sig.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaMemcpyAsync(w.tile_grp.p, T.grp.data(), T.grp.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaStreamSynchronize(h.stream));                      // T goes out of scope
  const size_t nv = (size_t)g.vlen();
  w.SL.ensure(nv * LT_ROWS_L); w.SU.ensure(nv * LT_ROWS_U);
  B200_CUDA(cudaMemsetAsync(w.SL.p, 0, nv * LT_ROWS_L * sizeof(double), h.stream));
  B200_CUDA(cudaMemsetAsync(w.SU.p, 0, nv * LT_ROWS_U * sizeof(double), h.stream));
  w.y.ensure(nv); w.x.ensure(nv);
  k_lane_sentinel<<<NUM_SMS * 8, 256, 0, h.stream>>>((long long)nv, w.x.p);
  B200_CUDA(cudaGetLastError());
  w.ready = true;
  if (getenv("B200_WAVE_DEBUG"))
    fprintf(stderr, "[lane] grid %d x %d x %d, %d planes per lane, %d strips x %d groups, %d tiles of %d steps, layout %.2f x rows, streams %.2f + %.2f GB\n", g.NR,
            g.NL, g.NP, g.TC, g.NS, g.NG, g.ntiles, g.NT, (double)nv / h.n, nv * LT_ROWS_L * 8e-9, nv * LT_ROWS_U * 8e-9);
}

void lane_refresh_values(Handle &h) {
  if (!h.lt.ready || h.n == 0) return;
  k_lane_fill<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, h.stream>>>(h.lt.g, h.lt.tile_of.p, h.n, h.d_rows.p, h.d_cols.p, h.d_ilu.p, h.lt.SL.p, h.lt.SU.p);
  B200_CUDA(cudaGetLastError());
}

template <bool UPPER, int TC>
static void lane_launch_tc(Handle &h, const double *S, double *out, double *r2) {
  const void *kern = (const void *)k_lane<UPPER, TC>;
  constexpr size_t SLOT = (size_t)TC * (UPPER ? LT_ROWS_U : LT_ROWS_L) * 256;
  int dev = 0, sms = 0, smem_max = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B200_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const int ntiles = h.lt.g.ntiles;
  int W = h.lt_warps > 0 ? h.lt_warps : (ntiles + sms - 1) / sms;   // all tiles co-resident when they fit
  W = std::max(1, std::min(7, W));                                 // + the producer warp = 256 threads: the whole register file for 8 warps
  int D = (int)(((size_t)smem_max - 1024) / ((size_t)W * (SLOT + 16)));
  if (h.lt_depth > 0) D = std::min(D, h.lt_depth);
  D = std::min(D, 16);
  B200_REQUIRE(D >= 2, "lane-tile triangular solve: ring does not fit in shared memory");
  const size_t smem = (size_t)W * D * (SLOT + 16);
  B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (W + 1) * 32, smem));
  B200_REQUIRE(per_sm >= 1, "lane-tile triangular solve: kernel does not fit on an SM");
  const int blocks = std::max(1, std::min(sms, (ntiles + W - 1) / W));
  LaneGeom g = h.lt.g; Ctrl *ctrl = h.ctrl.p; long long *trace = h.lt.trace_on ? h.lt.trace.p : nullptr;
  const int *tile_of = h.lt.tile_of.p, *tsig = h.lt.tile_sig.p, *tgrp = h.lt.tile_grp.p;
  int depth = D;
  void *argv[] = {(void *)&g, (void *)&tile_of, (void *)&tsig, (void *)&tgrp, (void *)&S, (void *)&out, (void *)&r2, (void *)&ctrl, (void *)&depth, (void *)&trace};
  B200_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3((W + 1) * 32), argv, smem, h.stream));
}

template <bool UPPER>
static void lane_launch(Handle &h, const double *S, double *out, double *r2) {
  switch (h.lt.g.TC) {
    case 1: lane_launch_tc<UPPER, 1>(h, S, out, r2); break;
    case 3: lane_launch_tc<UPPER, 3>(h, S, out, r2); break;
    case 4: lane_launch_tc<UPPER, 4>(h, S, out, r2); break;
    default: lane_launch_tc<UPPER, 2>(h, S, out, r2); break;
  }
}

void lu_apply_lane(Handle &h, double *u, const double *v) {
  B200_REQUIRE(h.lt.ready, "lane-tile triangular solve without a plan");
  LanePlan &w = h.lt;
  const int blocks = std::min((h.n + 255) / 256, NUM_SMS * 8);
  if (getenv("B200_LANE_TRACE") && !w.traced && h.st_pcond >= 2) { w.traced = true; lane_trace_enable(h, true); }   // the third application
  k_lane_in<<<blocks, 256, 0, h.stream>>>(w.g, w.tile_of.p, h.n, v, w.SL.p, w.g.vlen(), w.y.p);
  lane_launch<false>(h, w.SL.p, w.y.p, w.SU.p);
  lane_launch<true>(h, w.SU.p, w.x.p, nullptr);
  k_lane_out<<<blocks, 256, 0, h.stream>>>(w.g, w.tile_of.p, h.n, w.x.p, u);
  B200_CUDA(cudaGetLastError());
  h.st_launch += 4; h.st_pcond++;
  if (w.trace_on) {                                                // diagnostic: one traced application, written as text
    std::vector<long long> t; lane_trace_fetch(h, t);
    if (FILE *f = fopen(getenv("B200_LANE_TRACE"), "w")) {
      fprintf(f, "# sweep tile sigma group start_ns end_ns polls smid cyc_ring_wait cyc_poll first_ns cyc_shuffle cyc_request cyc_rows cyc_load cyc_resolve\n");
      std::vector<int> sg(w.g.ntiles), gr(w.g.ntiles);
      cudaMemcpy(sg.data(), w.tile_sig.p, sg.size() * sizeof(int), cudaMemcpyDeviceToHost); cudaMemcpy(gr.data(), w.tile_grp.p, gr.size() * sizeof(int), cudaMemcpyDeviceToHost);
      long long t0 = -1;
      for (size_t q = 0; q < t.size(); q += 16) if (t[q] && (t0 < 0 || t[q] < t0)) t0 = t[q];
      for (int sw = 0; sw < 2; ++sw) for (int k = 0; k < w.g.ntiles; ++k) {
        const long long *r = t.data() + ((size_t)sw * w.g.ntiles + k) * 16;
        fprintf(f, "%d %d %d %d %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", sw, k, sg[k], gr[k], r[0] - t0, r[1] - t0, r[2], r[3], r[4], r[5], r[6] ? r[6] - t0 : 0, r[8], r[9], r[10], r[11], r[12]);
      }
      fclose(f);
    }
    w.trace_on = false;
  }
}

// per-tile trace (B200_LANE_TRACE=file): 8 long long per (sweep, tile): start ns, end ns, polls, smid, cycles waiting for the ring, cycles polling, ns at which step 0 got its replayed values
void lane_trace_enable(Handle &h, bool on) {
  if (on) {
    const size_t m = (size_t)h.lt.g.ntiles * 2 * 16;
    h.lt.trace.ensure(m);
    B200_CUDA(cudaMemsetAsync(h.lt.trace.p, 0, m * sizeof(long long), h.stream));
  }
  h.lt.trace_on = on;
}
void lane_trace_fetch(Handle &h, std::vector<long long> &out) {
  out.assign((size_t)h.lt.g.ntiles * 2 * 16, 0);
  B200_CUDA(cudaStreamSynchronize(h.stream));
  B200_CUDA(cudaMemcpy(out.data(), h.lt.trace.p, out.size() * sizeof(long long), cudaMemcpyDeviceToHost));
}

}  // namespace b200

// Lane-tile triangular solves for the ILU(0) factor of a structured-grid stencil -- CRS_LUSolve, fem/src/CRSMatrix.F90:4590-4663.
// Geometry, data layout, shuffle routing and the row program: lanegeom.h (shared with the CPU emulation tests/lane_harness.cpp).
//
// Why: the dependency DAG of the 27-point factor has NR + 2 NL + 4 NP levels (1401 on the 200^3 heat problem) and every level-to-level
// hand-off that goes through L2 (level kernel: 0.9 us) or through shared memory + a CTA barrier (wave tiles: 0.3-0.55 us) sits on the
// critical path of the sweep.  Here a tile is ONE WARP: lane = sheared grid line, TC planes per lane, lanes skewed by two steps, so that
// every operand of a row is a register of the lane or arrives by one warp shuffle from lane j-1 (lanegeom.h).  A step (= one dependency
// level of the tile) is a few shuffles, TC x 13 multiply-subtract pairs in the reference's order and TC coalesced stores: no barrier, no
// shared-memory hand-off.  Only the lines just outside a tile (two ghost lanes, one plane below) come from L2, requested E steps ahead (template parameter, default 1)
// with the sentinel protocol; neighbouring tiles run concurrently, a few steps apart.
// Matrix entries + right-hand sides of a tile step are one contiguous block of the sweep's stream; every lane reads ITS column of it
// straight into registers two steps ahead (coalesced 256-byte warp loads, L1 bypassed) from L2, where lane 0 has put the block with
// cp.async.bulk.prefetch.L2 LT_PF steps earlier.  [Measured first: a shared-memory ring fed by bulk copies -- by the warp's own lane 0, then
// by a producer warp with full / empty mbarriers: 3.66 and 2.27 ms per application on the 200^3 problem against 2.02 ms now.]  The strong
// (L1-bypassing) loads of the replayed values come FIRST in a step: such a load does not issue before the warp's earlier loads have
// returned.  Arithmetic: the reference's operations in the reference's order, separate roundings; pad entries are (+0) x (+0).
// Bit-identical to the level kernel and to the CPU loop.  Which of the three kernels runs is decided per structure by timing them
// (precond.cu, tri_autotune_wave): flat grids go here, cubes to the wave tiles (profiles/r02_sptrsv_modes.txt).
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
// The same with both sides coalesced: a warp takes 32 consecutive steps of one plane slot of one tile, i.e. for each of the tile's 32 lanes
// 32 consecutive rows of its line (256 contiguous bytes of the natural-order vector), transposes them through shared memory and writes
// whole 32-wide rows of the layout (k_lane_in touches one 32-byte sector per 8-byte element on the layout side).
__global__ void __launch_bounds__(128) k_lane_in2(LaneGeom g, const int *__restrict__ tile_sig, const int *__restrict__ tile_grp, const double *__restrict__ v,
                                                  double *__restrict__ SL, double *__restrict__ y) {
  __shared__ double tr[4][32][33];
  const double sent = __longlong_as_double((long long)SENTINEL);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nch = g.NT / 32 + (g.NT % 32 ? 1 : 0);
  const long long nwork = (long long)g.ntiles * g.TC * nch, gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long q = gw; q < nwork; q += nwarps) {
    const int k = (int)(q / ((long long)g.TC * nch)), rem = (int)(q - (long long)k * g.TC * nch), p = rem / nch + 1, t0 = (rem % nch) * 32;
    const int sig = tile_sig[k], C = tile_grp[k];
    __syncwarp();
    for (int j = 0; j < 32; ++j) {                                  // natural -> shared: the warp's lanes are 32 consecutive rows of line j
      double val = 0.0;
      if (j >= LT_GH) {
        const LaneLine ln = lt_line(g, sig, C, j, p);
        const int a = t0 + lane - 2 * j - 2 * p;
        if (ln.valid && (unsigned)a < (unsigned)g.NR) val = v[a + (long long)g.NR * (ln.b + (long long)g.NL * ln.c)];
      }
      tr[wib][lane][j] = val;
    }
    __syncwarp();
    for (int r = 0; r < 32 && t0 + r < g.NT; ++r) {                 // shared -> layout rows
      const long long pos = (((long long)k * g.NT + t0 + r) * g.TC + (p - 1)) * 32 + lane;
      SL[lt_rhs_index(pos, LT_ROWS_L)] = tr[wib][r][lane];
      y[pos] = sent;
    }
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

__device__ __forceinline__ void lt_prefetch_l2(const void *p, unsigned bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory"); }
__device__ __forceinline__ void lt_prefetch_l2_if(const void *p, unsigned bytes, bool ok) {
  asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %2, 0;\n @q cp.async.bulk.prefetch.L2.global [%0], %1;\n }" ::"l"(p), "r"(bytes), "r"((int)ok) : "memory");
}
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
// High word of the sentinel: no value the sweeps store has it (NaN results are canonicalised before they are stored)
__device__ __forceinline__ bool lt_is_sentinel(double v) { return __double2hiint(v) == (int)(SENTINEL >> 32); }
__device__ __forceinline__ long long lt_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// ---- the sweep ------------------------------------------------------------------------------------------------------------------------
//   forward  (UPPER = false): out_i = rhs_i - sum_{j<i} L_ij out_j                   (4642-4649)
//   backward (UPPER = true) : out_i = Dinv_i * (rhs_i - sum_{j>i} U_ij out_j)       (4653-4660)
// S: this sweep's stream (entries + right-hand sides at pos(sweep coordinates)); Q: result at pos(mirrored sweep coordinates), pre-filled
// with the sentinel; R2: the backward stream, whose right-hand-side rows the forward sweep fills (nullptr for the backward sweep).
// Block = W warps, one tile each at a time; warp (block, w) takes tiles block + grid * w, + grid * W, ...: consecutive tiles
// (neighbouring start levels) sit on different SMs.
// Matrix entries: every lane reads ITS column of the tile step's block straight into registers (coalesced 256-byte warp loads, L1
// bypassed) LT_LEAD steps before use, from L2, where lane 0 has put the block LT_PF steps earlier with one cp.async.bulk.prefetch.L2
// per four steps.  [Measured first: a shared-memory ring fed by bulk copies, with a producer warp and a full / empty mbarrier pair per
// slot -- the consumer's mbarrier test + LDS + arrive cost 265 of 690 cycles per step.]
constexpr int LT_LEAD = 2;
template <bool UPPER, int TC, int E>
__global__ void __launch_bounds__(256, 1) k_lane(LaneGeom g, const int *__restrict__ tile_of, const int *__restrict__ tile_sig, const int *__restrict__ tile_grp,
                                                 const double *__restrict__ S, double *Q, double *R2, Ctrl *ctrl, long long *trace) {
  if (ctrl->done) return;
  constexpr int NE = UPPER ? 14 : 13, NROW = NE + 1;
  constexpr unsigned SLOT = TC * NROW * 256u;                       // bytes of a tile step
  constexpr long long BLKD = TC * NROW * 32;                        // doubles of a tile step
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int NT = g.NT, NR = g.NR, ntiles = g.ntiles, kstep = gridDim.x * W;
  constexpr long long STRIDE = TC * 32, RSTR = TC * LT_ROWS_U * 32;   // distance of consecutive rows of a line: result vector, backward stream
  constexpr long long RS2 = UPPER ? 1 : RSTR;                        // ... and in the second destination (natural order for the backward sweep)
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
      qp[p] = Q + q0 + (long long)(2 * lane + 2 * p) * STRIDE;
      // second destination of a result: forward sweep -> the right-hand-side row of the backward stream; backward sweep -> the
      // natural-order vector (row a of the mirrored line at base - a): no conversion pass afterwards
      if (!UPPER) rp[p] = R2 + lt_rhs_index(q0, LT_ROWS_U) + (long long)(2 * lane + 2 * p) * RSTR;
      else rp[p] = (R2 && ln.valid) ? R2 + (NR - 1) + (long long)NR * ((g.NL - 1 - ln.b) + (long long)g.NL * (g.NP - 1 - ln.c)) + (2 * lane + 2 * p) : nullptr;
    }
    int ab = -2 * lane;                                             // t0 - 2 lane (t0: first step of the unrolled group of 8)
    const double *sp = S + (long long)k * NT * BLKD + lane;          // the lane's column of the current step's block
    long long tr0 = 0, tr_polls = 0, tr_poll = 0, tr_first = 0, ph[5] = {0, 0, 0, 0, 0}, pc = 0;
    if (trace && lane == 0) tr0 = lt_gtime();
    if (lane == 0) lt_prefetch_l2(sp, (unsigned)(LT_PF < NT ? LT_PF : NT) * SLOT);
    LaneHist<TC> h;
    lt_hist_clear(h);
    double Hh[TC + 1][E + 1];
#pragma unroll
    for (int p = 0; p <= TC; ++p) {
#pragma unroll
      for (int u = 0; u < E; ++u) Hh[p][u] = lt_ld_relaxed_if(qp[p] - u * STRIDE, (unsigned)(ab + u - 2 * p) < nrq[p]);
      Hh[p][E] = 0.0;
    }
    double V[4][TC][NROW];                                          // entries + right-hand side of steps tau .. tau + LT_LEAD
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int p = 0; p < TC; ++p)
#pragma unroll
        for (int e = 0; e < NROW; ++e) V[q][p][e] = (q < LT_LEAD) ? ld_stream(sp + q * BLKD + (p * NROW + e) * 32) : 0.0;
    auto step = [&](auto Uc, int tau) {
      constexpr int U = decltype(Uc)::value;
      if (trace) pc = clock64();
      // replayed values of step tau + E.  FIRST: a strong (L1-bypassing) load does not issue before the warp's earlier loads have
      // returned (measured: 500-900 cycles behind the 15 streaming loads of a step, 70 when it comes before them)
#pragma unroll
      for (int p = 0; p <= TC; ++p) Hh[p][(U + E) % (E + 1)] = lt_ld_relaxed_if(qp[p] - (U + E) * STRIDE, (unsigned)(ab + U + E - 2 * p) < nrq[p]);
      // entries of step tau + LT_LEAD (the stream is padded by LT_LEAD steps: no guard at the end of the last tile)
#pragma unroll
      for (int p = 0; p < TC; ++p)
#pragma unroll
        for (int e = 0; e < NROW; ++e) V[(U + LT_LEAD) & 3][p][e] = ld_stream(sp + (U + LT_LEAD) * BLKD + (p * NROW + e) * 32);
      if ((U & 3) == 0) lt_prefetch_l2_if(sp + (long long)(U + LT_PF) * BLKD, 4u * SLOT, lane == 0 && tau + LT_PF + 4 <= NT);
      if (trace) { const long long c = clock64(); ph[3] += c - pc; pc = c; }
      // the two shuffles
      LaneMsg<TC> m;
      lt_send<TC, U>(h, m);
#pragma unroll
      for (int p = 0; p <= TC; ++p) m.r[p] = __shfl_up_sync(FULL, m.r[p], 1);
#pragma unroll
      for (int p = 0; p < TC; ++p) m.t[p] = __shfl_up_sync(FULL, m.t[p], 1);
      lt_recv<TC, U>(h, m);
      if (trace) { const long long c = clock64(); ph[0] += c - pc; pc = c; }
      if (trace) { const long long c = clock64(); ph[1] += c - pc; pc = c; }
      // rows
      double out[TC + 1];
      out[0] = 0.0;
#pragma unroll
      for (int p = 1; p <= TC; ++p) {
        double acc = lt_row<UPPER, TC, U>(h, p, V[U & 3][p - 1], V[U & 3][p - 1][NE]);
        if (acc != acc) acc = __longlong_as_double((long long)CANON_NAN);
        const bool active = (unsigned)(ab + U - 2 * p) < nrs[p];
        if (!active) acc = 0.0;
        out[p] = acc;
        lt_st_relaxed_if(qp[p] - U * STRIDE, acc, active);
        lt_st_if(rp[p] - U * RS2, acc, active && rp[p] != nullptr);
      }
      if (trace) { const long long c = clock64(); ph[2] += c - pc; pc = c; }
      // replayed values of this step (requested E steps ago; a producer that is not that far ahead yet is polled)
      double hv[TC + 1];
      bool need = false;
#pragma unroll
      for (int p = 0; p <= TC; ++p) { hv[p] = Hh[p][U % (E + 1)]; need = need || lt_is_sentinel(hv[p]); }
      if (__any_sync(FULL, need)) {
        const long long c0 = trace ? clock64() : 0;
        unsigned tries = 0;
        do {
          need = false;
#pragma unroll
          for (int p = 0; p <= TC; ++p)
            if (lt_is_sentinel(hv[p])) { hv[p] = lt_ld_relaxed(qp[p] - U * STRIDE); need = need || lt_is_sentinel(hv[p]); ++tr_polls; }
          if (need) {
            if (++spins > LT_SPIN_LIMIT) {
              ctrl->spin_timeout = 1;
#pragma unroll
              for (int p = 0; p <= TC; ++p) if (lt_is_sentinel(hv[p])) hv[p] = 0.0;
              need = false;
            } else if (++tries > 8) __nanosleep(tries > 64 ? 400 : 100);
          }
        } while (__any_sync(FULL, need));
        if (trace) { tr_poll += clock64() - c0; if (tau == 0) tr_first = lt_gtime(); }
      }
#pragma unroll
      for (int p = 0; p <= TC; ++p) {
        h.X[p][U & 7] = (p == 0 || lane < LT_GH) ? hv[p] : out[p];
        if (U == 7) { qp[p] -= 8 * STRIDE; if (rp[p]) rp[p] -= 8 * RS2; }
      }
      if (U == 7) { ab += 8; sp += 8 * BLKD; }
      if (trace) { const long long c = clock64(); ph[4] += c - pc; pc = c; }
    };
    static_assert(8 % (E + 1) == 0, "replay ring must divide the unroll factor");
    for (int t0 = 0; t0 < NT; t0 += 8) {
      step(std::integral_constant<int, 0>{}, t0); step(std::integral_constant<int, 1>{}, t0 + 1);
      step(std::integral_constant<int, 2>{}, t0 + 2); step(std::integral_constant<int, 3>{}, t0 + 3);
      step(std::integral_constant<int, 4>{}, t0 + 4); step(std::integral_constant<int, 5>{}, t0 + 5);
      step(std::integral_constant<int, 6>{}, t0 + 6); step(std::integral_constant<int, 7>{}, t0 + 7);
    }
    if (trace) {
      long long *r = trace + ((UPPER ? g.ntiles : 0) + (long long)k) * 16;
      if (lane == 0) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); r[0] = tr0; r[1] = lt_gtime(); r[3] = smid; r[5] = tr_poll; r[6] = tr_first; for (int q = 0; q < 5; ++q) r[8 + q] = ph[q]; }
      if (tr_polls) atomicAdd((unsigned long long *)(r + 2), (unsigned long long)tr_polls);
    }
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------------------------
void lane_release(Handle &h) {
  LanePlan &w = h.lt;
  w.SL.release(); w.SU.release(); w.y.release(); w.x.release(); w.tile_of.release(); w.tile_sig.release(); w.tile_grp.release(); w.trace.release(); w.mapL.release(); w.mapU.release();
  w.ready = false; w.tried = false;
}

void lane_analyse(Handle &h) {
  LanePlan &w = h.lt;
  if (w.ready || w.tried) return;
  w.tried = true;
  const char *why = nullptr;
  SkewGeom sg;
  if (h.ilu_sep()) why = "ILU(n > 0) / BILU pattern";
  else why = sk_detect(h.n, h.h_rows.data(), h.h_cols.data(), h.h_diag.data(), sg);
  if (why) {
    if (getenv("B200_WAVE_DEBUG")) fprintf(stderr, "[lane] not usable (%s): level kernel stays\n", why);
    return;
  }
  const int TC = std::max(1, std::min(2, h.lt_tc));
  LaneTiles T;
  lt_plan(w.g, sg.NR, sg.NL, sg.NP, TC, T);
  const LaneGeom &g = w.g;
  w.tile_of.ensure(T.tile_of.size()); w.tile_sig.ensure(T.sig.size()); w.tile_grp.ensure(T.grp.size());
  B200_CUDA(cudaMemcpyAsync(w.tile_of.p, T.tile_of.data(), T.tile_of.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaMemcpyAsync(w.tile_sig.p, T.sig.data(), T.sig.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaMemcpyAsync(w.tile_grp.p, T.grp.data(), T.grp.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaStreamSynchronize(h.stream));                      // T goes out of scope
  const size_t nv = (size_t)g.vlen();
  const size_t pad = (size_t)4 * g.TC * 32;                          // the sweeps read LT_LEAD steps past the last tile
  w.SL.ensure((nv + pad) * LT_ROWS_L); w.SU.ensure((nv + pad) * LT_ROWS_U);
  B200_CUDA(cudaMemsetAsync(w.SL.p, 0, (nv + pad) * LT_ROWS_L * sizeof(double), h.stream));
  B200_CUDA(cudaMemsetAsync(w.SU.p, 0, (nv + pad) * LT_ROWS_U * sizeof(double), h.stream));
  {                                                                 // refill map (structure.cu): the scatter once, on indices
    DBuf<double> idx; idx.ensure((size_t)h.nnz);
    stream_iota1(h, h.nnz, idx.p);
    k_lane_fill<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, h.stream>>>(g, w.tile_of.p, h.n, h.d_rows.p, h.d_cols.p, idx.p, w.SL.p, w.SU.p);
    w.mapL.ensure((nv + pad) * LT_ROWS_L); w.mapU.ensure((nv + pad) * LT_ROWS_U);
    stream_map_build(h, (long long)((nv + pad) * LT_ROWS_L), w.SL.p, w.mapL.p);
    stream_map_build(h, (long long)((nv + pad) * LT_ROWS_U), w.SU.p, w.mapU.p);
    B200_CUDA(cudaStreamSynchronize(h.stream));
    idx.release();
  }
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
  const size_t nv = (size_t)h.lt.g.vlen(), pad = (size_t)4 * h.lt.g.TC * 32;
  stream_gather(h, (long long)((nv + pad) * LT_ROWS_L), h.lt.mapL.p, h.d_ilu.p, h.lt.SL.p);     // (right-hand-side rows: map -1 -> 0; every application rewrites them)
  stream_gather(h, (long long)((nv + pad) * LT_ROWS_U), h.lt.mapU.p, h.d_ilu.p, h.lt.SU.p);
}

template <bool UPPER, int TC>
static void lane_launch_tc(Handle &h, const double *S, double *out, double *r2) {
  const void *kern = h.lt_e == 1 ? (const void *)k_lane<UPPER, TC, 1> : (h.lt_e == 7 ? (const void *)k_lane<UPPER, TC, 7> : (const void *)k_lane<UPPER, TC, 3>);
  int dev = 0, sms = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ntiles = h.lt.g.ntiles;
  int W = h.lt_warps > 0 ? h.lt_warps : (ntiles + sms - 1) / sms;   // all tiles co-resident when they fit
  W = std::max(1, std::min(8, W));
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, W * 32, 0));
  B200_REQUIRE(per_sm >= 1, "lane-tile triangular solve: kernel does not fit on an SM");
  const int blocks = std::max(1, std::min(sms, (ntiles + W - 1) / W));
  LaneGeom g = h.lt.g; Ctrl *ctrl = h.ctrl.p; long long *trace = h.lt.trace_on ? h.lt.trace.p : nullptr;
  const int *tile_of = h.lt.tile_of.p, *tsig = h.lt.tile_sig.p, *tgrp = h.lt.tile_grp.p;
  void *argv[] = {(void *)&g, (void *)&tile_of, (void *)&tsig, (void *)&tgrp, (void *)&S, (void *)&out, (void *)&r2, (void *)&ctrl, (void *)&trace};
  B200_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(W * 32), argv, 0, h.stream));
}

template <bool UPPER>
static void lane_launch(Handle &h, const double *S, double *out, double *r2) {
  switch (h.lt.g.TC) {
    case 2: lane_launch_tc<UPPER, 2>(h, S, out, r2); break;
    default: lane_launch_tc<UPPER, 1>(h, S, out, r2); break;
  }
}

void lu_apply_lane(Handle &h, double *u, const double *v) {
  B200_REQUIRE(h.lt.ready, "lane-tile triangular solve without a plan");
  LanePlan &w = h.lt;
  const int blocks = std::min((h.n + 255) / 256, NUM_SMS * 8);
  if (getenv("B200_LANE_TRACE") && !w.traced && h.st_pcond >= 2) { w.traced = true; lane_trace_enable(h, true); }   // the third application
  static const bool conv2 = !(getenv("B200_LANE_CONV") && atoi(getenv("B200_LANE_CONV")) == 1);      // 1: element-wise k_lane_in
  if (conv2) k_lane_in2<<<NUM_SMS * 12, 128, 0, h.stream>>>(w.g, w.tile_sig.p, w.tile_grp.p, v, w.SL.p, w.y.p);
  else k_lane_in<<<blocks, 256, 0, h.stream>>>(w.g, w.tile_of.p, h.n, v, w.SL.p, w.g.vlen(), w.y.p);
  lane_launch<false>(h, w.SL.p, w.y.p, w.SU.p);
  static const bool direct = !(getenv("B200_LANE_OUT") && atoi(getenv("B200_LANE_OUT")) == 1);      // 1: conversion pass k_lane_out
  lane_launch<true>(h, w.SU.p, w.x.p, direct ? u : nullptr);
  if (direct) k_lane_sentinel<<<NUM_SMS * 8, 256, 0, h.stream>>>(w.g.vlen(), w.x.p);   // the slots are handed back as sentinels
  else k_lane_out<<<blocks, 256, 0, h.stream>>>(w.g, w.tile_of.p, h.n, w.x.p, u);
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

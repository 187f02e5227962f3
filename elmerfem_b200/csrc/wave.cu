// Wave-tile triangular solves for the ILU(0) factor of a structured-grid stencil -- CRS_LUSolve, fem/src/CRSMatrix.F90:4590-4663.
// Geometry, data layout and operand routing: wavegeom.h (shared with the CPU emulation tests/wave_harness.cpp).
//
// Why: the level kernel pays one L2 hand-off (0.43 us store -> poll, plus the row itself) per dependency level, 2 x 1401 of them on the
// 200^3 heat problem.  Here a CTA owns a TILE of TB x TC grid lines (one thread per line) that advance in lockstep, one row per thread
// and step: all hand-offs inside a tile go through a shared-memory ring and one CTA barrier per step; only the TB + 2 + 2 TC halo lines
// of a tile come from L2, fetched E steps ahead by loader threads (sentinel protocol: the result vector is pre-filled with a NaN payload
// no arithmetic produces, a loader that still finds it polls).  Shearing the line coordinate (beta = b + c) makes the tile DAG acyclic
// with dependencies only towards smaller (sigma, C), so neighbouring tiles run concurrently, one L2 hop apart, and the critical path
// holds (number of tile rows + columns) hops instead of one per level.  Matrix entries and right-hand sides of a tile step are
// contiguous rows of NTHR values; every compute thread reads ITS column of them straight into registers three to five steps ahead
// (coalesced 256-byte warp loads, L1 bypassed) from L2, where a prefetch warp has put the block with cp.async.bulk.prefetch.L2; steps in
// which the thread has no row are not loaded.  [Measured first and dropped: a TMA bulk-copy ring (one thread pays ~300 cycles per step for
// the expect_tx + copy issue and ~260 for the mbarrier wait) and a cp.async ring (+240 cycles of LSU issue per step).]  The backward sweep
// stores its result in natural order as well, so only the way in needs a layout pass (k_wave_in2).
// Arithmetic: the reference's operations in the reference's order (entries in ascending column order, separate multiply / subtract
// roundings, inverse diagonal last); pad entries (neighbours outside the grid) are (+0) x (+0).  Bit-identical to the level kernel and
// to the CPU loop.
#include "common.cuh"
#include "kernels.cuh"
#include "wavegeom.h"
#include <algorithm>

namespace b200 {

constexpr long long WV_SPIN_LIMIT = 1LL << 22;

// ---- values: CRS order of the ILU factor -> the two per-step streams ---------------------------------------------------------------
__global__ void k_wave_fill(WaveGeom g, const int *__restrict__ tile_of, int n, const int *__restrict__ rows, const int *__restrict__ cols,
                            const double *__restrict__ ilu, double *__restrict__ SL, double *__restrict__ SU) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) wv_fill_row(g, tile_of, i, rows, cols, ilu, SL, SU);
}
// natural order -> tile layout (right-hand side of the forward sweep) + sentinel fill of the forward result
__global__ void k_wave_in(WaveGeom g, const int *__restrict__ tile_of, int n, const double *__restrict__ v, double *__restrict__ yin, long long nv,
                          double *__restrict__ y) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = t0; i < nv; i += stride) y[i] = sent;
  for (long long i = t0; i < n; i += stride) {
    const int a = (int)(i % g.NR), b = (int)((i / g.NR) % g.NL), c = (int)(i / ((long long)g.NR * g.NL));
    yin[wv_pos(g, tile_of, a, b, c)] = v[i];
  }
}
// tile layout -> natural order; the slots are handed back as sentinels for the next application
__global__ void k_wave_out(WaveGeom g, const int *__restrict__ tile_of, int n, double *__restrict__ x, double *__restrict__ u) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(i % g.NR), b = (int)((i / g.NR) % g.NL), c = (int)(i / ((long long)g.NR * g.NL));
    const long long p = wv_pos(g, tile_of, a, b, c);
    u[i] = x[p]; x[p] = sent;
  }
}
__global__ void k_wave_sentinel(long long nv, double *__restrict__ x) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) x[i] = sent;
}

// The way in with BOTH sides coalesced: a block takes a chunk of 32 consecutive steps of one tile, i.e. for every line of
// the tile 32 consecutive rows a.  Natural side: one warp per line, lanes = the 32 rows (256 contiguous bytes); tile side: rows of NTHR
// values per step; the skewed transposition goes through shared memory (row length NTHR + 1: conflict-free).  k_wave_in
// touches one 32-byte sector per 8-byte element on the tile side (115 us per application on the 200^3 problem).
constexpr int WV_CH = 32;
__global__ void __launch_bounds__(256) k_wave_in2(WaveGeom g, const int *__restrict__ tile_sig, const int *__restrict__ tile_grp, const double *__restrict__ v,
                                                  double *__restrict__ yin, double *__restrict__ y) {
  extern __shared__ double wv_tr[];                                 // [WV_CH][NTHR + 1]
  const double sent = __longlong_as_double((long long)SENTINEL);
  const int NTHR = g.nthr(), LD = NTHR + 1, nch = (g.NT + WV_CH - 1) / WV_CH;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (long long q = blockIdx.x; q < (long long)g.ntiles * nch; q += gridDim.x) {
    const int k = (int)(q / nch), t0 = (int)(q - (long long)k * nch) * WV_CH;
    const int sig = tile_sig[k], C = tile_grp[k];
    __syncthreads();                                                // the previous chunk has been written out
    for (int t = wib; t < NTHR; t += nw) {                          // natural -> shared: warp per line
      const WaveLine ln = wv_line(g, sig, C, t % g.TB, t / g.TB);
      const int a = t0 + lane - ln.tau0;
      double val = 0.0;
      if (ln.valid && (unsigned)a < (unsigned)g.NR && t0 + lane < g.NT) val = v[a + (long long)g.NR * (ln.b + (long long)g.NL * ln.c)];
      wv_tr[lane * LD + t] = val;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < WV_CH * NTHR; e += blockDim.x) {  // shared -> tile layout, whole step rows
      const int s = e / NTHR, t = e - s * NTHR;
      if (t0 + s < g.NT) {
        const long long p = ((long long)k * g.NT + t0 + s) * NTHR + t;
        yin[p] = wv_tr[s * LD + t]; y[p] = sent;
      }
    }
  }
}
__device__ __forceinline__ unsigned wv_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wv_prefetch_l2(const void *p, unsigned bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory"); }
#ifdef WV_NO_BAR
__device__ __forceinline__ void wv_bar(int nthreads) { asm volatile("" ::: "memory"); }
#else
__device__ __forceinline__ void wv_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
#endif
// the sweep's own store: no compiler barrier (nothing in this thread reads the result vector)
#ifndef WV_ST
#define WV_ST 0
#endif
__device__ __forceinline__ void wv_st_relaxed(double *p, double v) {
#if WV_ST == 0
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v));
#elif WV_ST == 1
  asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v));
#elif WV_ST == 2
  asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v));
#endif
}
__device__ __forceinline__ double wv_ld_relaxed(const double *p) {
  double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ long long wv_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// ---- the sweep ------------------------------------------------------------------------------------------------------------------------
//   forward  (UPPER = false): out_i = rhs_i - sum_{j<i} L_ij out_j                   (4642-4649)
//   backward (UPPER = true) : out_i = Dinv_i * (rhs_i - sum_{j>i} U_ij out_j)       (4653-4660)
// S: matrix stream, RHS: right-hand side at pos(sweep coordinates); Q: result at pos(mirrored sweep coordinates), pre-filled with the
// sentinel.  E: steps a halo row is requested ahead; PF: steps the L2 prefetch runs ahead.
// Block = TB * TC compute threads + HW loader warps (these take part in the step barrier) + one free-running prefetch warp.
//
// Matrix entries: every thread reads ITS column of the stream straight into registers (coalesced 256-byte warp loads, L1 bypassed),
// three steps before the value is used, from L2 -- where the prefetch warp has put the block PF steps earlier with one
// cp.async.bulk.prefetch.L2 per WV_PFG steps.  Steps in which the thread has no row are not loaded.  [Measured alternatives: a TMA
// bulk-copy ring costs one thread ~300 cycles per step for expect_tx + copy issue and ~260 for the mbarrier wait; a cp.async ring
// costs the LSU 16 LDGSTS + 15 LDS per thread and step, +240 cycles per step.]
constexpr int WV_PFG = 4;      // steps per prefetch instruction
template <bool UPPER, int TB, int TC, int E, int PF>
__global__ void __launch_bounds__(TB * TC + 32 * ((TB + 2 + 2 * TC + 31) / 32) + 32) k_wave(WaveGeom g, const int *__restrict__ tile_of, const int *__restrict__ tile_sig,
                                                                                            const int *__restrict__ tile_grp, const double *__restrict__ S,
                                                                                            const double *__restrict__ RHS, double *Q, Ctrl *ctrl, long long *trace, double *nat) {
  if (ctrl->done) return;
  constexpr int NTHR = TB * TC, NH = TB + 2 + 2 * TC, NE = UPPER ? 14 : 13;
  constexpr int YW = TB + 2, YH = TC + 1, YSLOT = YW * YH;
  constexpr int NALL = NTHR + 32 * ((NH + 31) / 32);                // threads of the step barrier
  constexpr int NA = UPPER ? 7 : 5, NB = 2, NC = UPPER ? 6 : 7;     // values per step of the three chains (see below)
  extern __shared__ __align__(128) unsigned char wv_smem[];
  double *Yr = reinterpret_cast<double *>(wv_smem);
  volatile int *done_steps = reinterpret_cast<volatile int *>(Yr + WV_RING * YSLOT);   // steps of the current tile behind the barrier
  const int tid = threadIdx.x;
  const int NT = g.NT, NR = g.NR;
  long long spins = 0;
  for (int k = blockIdx.x; k < g.ntiles; k += gridDim.x) {
    const int sig = tile_sig[k], C = tile_grp[k];
    const long long step0 = (long long)k * NT;
    __syncthreads();                                                // every warp (the free-running prefetch warp too) has left the previous tile
    for (int q = tid; q < WV_RING * YSLOT; q += NALL + 32) Yr[q] = 0.0;
    if (tid == 0) *done_steps = 0;
    long long tr0 = 0, tr_polls = 0;
    if (trace && tid == NTHR) tr0 = wv_gtime();
    __syncthreads();                                                // ring zeroed
    if (tid < NTHR) {
      // ---------------- compute thread: line (jb, w) ----------------
      // Three independent chains per step, so that only the five terms that need the results of step tau - 1 follow the barrier:
      //   A: row of step tau      terms 8..12 (B: a+1 | A: a-1, a, a+1 | own a-1) on `hi`, publish               (backward: the whole row)
      //   B: row of step tau + 1  terms 6, 7  (B: a-1, a)                          mid -> hi
      //   C: row of step tau + 2  terms 0..5  (D, C: a-1, a, a+1)                  rhs -> mid
      // in the reference's left-to-right order (4642-4649).  The backward row starts with its newest operand (4653-4660): nothing of it
      // can be summed early, only the products are formed ahead.
      // Registers: VA / VB / VC hold the entries chain A / B / C needs, four steps deep (loop unrolled by four: constant indices);
      // at step tau the loads of steps tau + 3 (A), tau + 4 (B), tau + 5 (C) are issued, i.e. each value three steps before its use.
      const int jb = tid % TB, w = tid / TB;
      const WaveLine ln = wv_line(g, sig, C, jb, w);
      double *qp = Q + (ln.valid ? wv_pos_mirror(g, tile_of, 0, ln.b, ln.c) : 0);     // row a at qp - a * NTHR
      // the backward sweep also stores its result in natural order (row a of the mirrored line at np - a): no conversion pass afterwards
      double *np = (nat && ln.valid) ? nat + (g.NR - 1) + (long long)g.NR * ((g.NL - 1 - ln.b) + (long long)g.NL * (g.NP - 1 - ln.c)) : nullptr;
      const double *yA = Yr + (w + 1) * YW + (jb + 1), *yB = Yr + w * YW + (jb + 2), *yC = Yr + w * YW + (jb + 1), *yD = Yr + w * YW + jb;
      double *yO = Yr + (w + 1) * YW + (jb + 2);
      // this thread's column of the stream / right-hand side at the first step of the unrolled group of four: the loads of a step use
      // compile-time offsets from it (one pointer update per four steps instead of a 64-bit multiply per load)
      const double *gS = S + step0 * (NE * NTHR) + tid, *gR = RHS + step0 * NTHR + tid;
      const int tau0 = ln.tau0;
      const bool valid = ln.valid;
      double *qrow = qp + (long long)tau0 * NTHR;                    // row of step t0 (first of the group of four) at qrow; a = tau - tau0
      double *nrow = np ? np + tau0 : nullptr;
      double Am = 0.0, A0 = 0.0, B0 = 0.0, Cm = 0.0, C0 = 0.0, Dm = 0.0, D0 = 0.0, h = 0.0;
      double hi = 0.0, mid = 0.0;                                   // forward: partial rows of steps tau and tau + 1
      double P1[8], P2[6];                                          // backward: products of steps tau (terms 0..7) and tau + 1 (terms 0..5)
      double VA[4][NA], VB[4][NB], VC[4][NC];
#pragma unroll
      for (int e = 0; e < 8; ++e) P1[e] = 0.0;
#pragma unroll
      for (int e = 0; e < 6; ++e) P2[e] = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int e = 0; e < NA; ++e) VA[q][e] = 0.0;
#pragma unroll
        for (int e = 0; e < NB; ++e) VB[q][e] = 0.0;
#pragma unroll
        for (int e = 0; e < NC; ++e) VC[q][e] = 0.0;
      }
      // (no row of a compute thread lies in steps 0 .. WV_PRE - 1, so nothing has to be loaded before the loop)
      for (int t0 = 0; t0 < NT; t0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int tau = t0 + u;
          if (tau < NT) {
            const int un = (u + 3) & 3;
            {                                                       // chain A of step tau + 3: terms 8..12 (+ inverse diagonal, right-hand side)
              const int st = tau + 3;
              if (valid && (unsigned)(st - tau0) < (unsigned)NR) {
                const double *src = gS + (u + 3) * (NE * NTHR);
#pragma unroll
                for (int e = 0; e < 5; ++e) VA[un][e] = ld_stream(src + (8 + e) * NTHR);
                if (UPPER) { VA[un][5] = ld_stream(src + 13 * NTHR); VA[un][6] = ld_stream(gR + (u + 3) * NTHR); }
              }
            }
            {                                                       // chain B of step tau + 4: terms 6, 7
              const int st = tau + 4;
              if (valid && (unsigned)(st - tau0) < (unsigned)NR) {
                const double *src = gS + (u + 4) * (NE * NTHR);
                VB[un][0] = ld_stream(src + 6 * NTHR); VB[un][1] = ld_stream(src + 7 * NTHR);
              }
            }
            {                                                       // chain C of step tau + 5: terms 0..5 (+ right-hand side)
              const int st = tau + 5;
              if (valid && (unsigned)(st - tau0) < (unsigned)NR) {
                const double *src = gS + (u + 5) * (NE * NTHR);
#pragma unroll
                for (int e = 0; e < 6; ++e) VC[un][e] = ld_stream(src + e * NTHR);
                if (!UPPER) VC[un][6] = ld_stream(gR + (u + 5) * NTHR);
              }
            }
            const int r1 = ((tau - 1) & (WV_RING - 1)) * YSLOT, r3 = ((tau - 3) & (WV_RING - 1)) * YSLOT;
            const double An = yA[r1], Bn = yB[r1], Cn = yC[r1], Dn = yD[r3];
            const int a = tau - tau0;
            const bool active = valid && (unsigned)a < (unsigned)NR;
            double acc;
            if (!UPPER) {
              acc = nfms(hi, VA[u][0], Bn); acc = nfms(acc, VA[u][1], Am); acc = nfms(acc, VA[u][2], A0);
              acc = nfms(acc, VA[u][3], An); acc = nfms(acc, VA[u][4], h);
              hi = nfms(mid, VB[u][0], B0); hi = nfms(hi, VB[u][1], Bn);
              mid = nfms(VC[u][6], VC[u][0], Dm); mid = nfms(mid, VC[u][1], D0); mid = nfms(mid, VC[u][2], Dn);
              mid = nfms(mid, VC[u][3], Cm); mid = nfms(mid, VC[u][4], C0); mid = nfms(mid, VC[u][5], Cn);
            } else {
              acc = nfms(VA[u][6], VA[u][4], h); acc = nfms(acc, VA[u][3], An); acc = nfms(acc, VA[u][2], A0);
              acc = nfms(acc, VA[u][1], Am); acc = nfms(acc, VA[u][0], Bn);
#pragma unroll
              for (int e = 7; e >= 0; --e) acc = __dsub_rn(acc, P1[e]);
              acc = __dmul_rn(VA[u][5], acc);
#pragma unroll
              for (int e = 0; e < 6; ++e) P1[e] = P2[e];
              P1[6] = __dmul_rn(VB[u][0], B0); P1[7] = __dmul_rn(VB[u][1], Bn);
              P2[0] = __dmul_rn(VC[u][0], Dm); P2[1] = __dmul_rn(VC[u][1], D0); P2[2] = __dmul_rn(VC[u][2], Dn);
              P2[3] = __dmul_rn(VC[u][3], Cm); P2[4] = __dmul_rn(VC[u][4], C0); P2[5] = __dmul_rn(VC[u][5], Cn);
            }
            if (acc != acc) acc = __longlong_as_double((long long)CANON_NAN);
            if (!active) acc = 0.0;
            yO[(tau & (WV_RING - 1)) * YSLOT] = acc;
            if (active) { wv_st_relaxed(qrow - u * NTHR, acc); if (UPPER && np) nrow[-u] = acc; }
            h = acc;
            Am = A0; A0 = An; B0 = Bn; Cm = C0; C0 = Cn; Dm = D0; D0 = Dn;
            wv_bar(NALL);
          }
        }
        gS += 4 * (NE * NTHR); gR += 4 * NTHR; qrow -= 4 * NTHR; nrow -= 4;
      }
    } else if (tid < NALL) {
      // ---------------- loader thread: halo line hh ----------------
      const int hh = tid - NTHR;
      int jb = 0, w = 0;
      bool on = hh < NH;
      if (on) wv_halo(g, hh, jb, w);
      const WaveLine ln = wv_line(g, sig, C, jb, w);
      on = on && ln.valid;
      const double *qp = Q + (on ? wv_pos_mirror(g, tile_of, 0, ln.b, ln.c) : 0);
      double *yO = Yr + (w + 1) * YW + (jb + 2);
      const int tau0 = ln.tau0;
      double H[E + 1];
      long long cy_bar = 0;
#pragma unroll
      for (int q = 0; q < E + 1; ++q) H[q] = 0.0;
      auto request = [&](int tau) -> double {                     // the row the line publishes at step tau
        const int a = tau - tau0;
        double r = 0.0;
        if (on && (unsigned)a < (unsigned)NR) r = wv_ld_relaxed(qp - (long long)a * NTHR);
        return r;
      };
#pragma unroll
      for (int u = 0; u < E; ++u) H[u] = request(u);               // steps 0 .. E-1
      for (int t0 = 0; t0 < NT; t0 += E + 1) {
#pragma unroll
        for (int u = 0; u < E + 1; ++u) {
          const int tau = t0 + u;
          if (tau < NT) {
            H[(u + E) % (E + 1)] = request(tau + E);
            double head = H[u];
            if (is_sentinel(head)) {
              const double *pa = qp - (long long)(tau - tau0) * NTHR;
              do {
                if (++spins > WV_SPIN_LIMIT) { ctrl->spin_timeout = 1; head = 0.0; break; }
                ++tr_polls;
                head = wv_ld_relaxed(pa);
              } while (is_sentinel(head));
            }
            if (hh < NH) yO[(tau & (WV_RING - 1)) * YSLOT] = head;
            const long long c1 = trace ? clock64() : 0;
            wv_bar(NALL);
            if (trace) cy_bar += clock64() - c1;
            if (tid == NTHR) *done_steps = tau + 1;
          }
        }
      }
      if (trace) {
        if (tid == NTHR) { long long *r = trace + ((UPPER ? g.ntiles : 0) + (long long)k) * 8; unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); r[0] = tr0; r[1] = wv_gtime(); r[3] = smid; r[5] = cy_bar; }
        if (tr_polls) atomicAdd((unsigned long long *)(trace + ((UPPER ? g.ntiles : 0) + (long long)k) * 8 + 2), (unsigned long long)tr_polls);
      }
    } else if (PF > 0 && tid == NALL) {
      // ---------------- prefetch warp (one lane): stream blocks of WV_PFG steps into L2, at most PF steps ahead of the barrier ----------------
      for (int st = 0; st < NT; st += WV_PFG) {
        while (*done_steps + PF < st) { if (++spins > WV_SPIN_LIMIT) break; }
        const int ns = (NT - st < WV_PFG) ? NT - st : WV_PFG;
        wv_prefetch_l2(S + (step0 + st) * (NE * NTHR), (unsigned)(ns * NE * NTHR * 8));
        wv_prefetch_l2(RHS + (step0 + st) * NTHR, (unsigned)(ns * NTHR * 8));
      }
    }
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------------------------
void wave_release(Handle &h) {
  WavePlan &w = h.wv;
  w.SL.release(); w.SU.release(); w.yin.release(); w.y.release(); w.x.release(); w.tile_of.release(); w.tile_sig.release(); w.tile_grp.release(); w.trace.release(); w.mapL.release(); w.mapU.release();
  w.ready = false; w.tried = false;
}

void wave_analyse(Handle &h) {
  WavePlan &w = h.wv;
  if (w.ready || w.tried) return;
  w.tried = true;
  const char *why = nullptr;
  SkewGeom sg;
  if (h.ilu_sep()) why = "ILU(n > 0) / BILU pattern";
  else why = sk_detect(h.n, h.h_rows.data(), h.h_cols.data(), h.h_diag.data(), sg);
  if (why) {
    if (getenv("B200_WAVE_DEBUG")) fprintf(stderr, "[wave] not usable (%s): level kernel stays\n", why);
    return;
  }
  const int cfg = h.wv_cfg;                                         // tile shape; the ring depth / request lead go with it (wave_launch)
  const int TB = (cfg == 5) ? 32 : ((cfg == 3) ? 8 : 16), TC = (cfg == 1 || cfg == 5) ? 4 : 8;
  WaveTiles T;
  wv_plan(w.g, sg.NR, sg.NL, sg.NP, TB, TC, T);
  const WaveGeom &g = w.g;
  w.tile_of.ensure(T.tile_of.size()); w.tile_sig.ensure(T.sig.size()); w.tile_grp.ensure(T.grp.size());
  B200_CUDA(cudaMemcpyAsync(w.tile_of.p, T.tile_of.data(), T.tile_of.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaMemcpyAsync(w.tile_sig.p, T.sig.data(), T.sig.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaMemcpyAsync(w.tile_grp.p, T.grp.data(), T.grp.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaStreamSynchronize(h.stream));                      // T goes out of scope
  const size_t steps = (size_t)g.nsteps(), nthr = (size_t)g.nthr();
  w.SL.ensure(steps * 13 * nthr); w.SU.ensure(steps * 14 * nthr);
  B200_CUDA(cudaMemsetAsync(w.SL.p, 0, steps * 13 * nthr * sizeof(double), h.stream));
  B200_CUDA(cudaMemsetAsync(w.SU.p, 0, steps * 14 * nthr * sizeof(double), h.stream));
  {                                                                 // refill map (structure.cu): the scatter once, on indices
    DBuf<double> idx; idx.ensure((size_t)h.nnz);
    stream_iota1(h, h.nnz, idx.p);
    k_wave_fill<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, h.stream>>>(g, w.tile_of.p, h.n, h.d_rows.p, h.d_cols.p, idx.p, w.SL.p, w.SU.p);
    w.mapL.ensure(steps * 13 * nthr); w.mapU.ensure(steps * 14 * nthr);
    stream_map_build(h, (long long)(steps * 13 * nthr), w.SL.p, w.mapL.p);
    stream_map_build(h, (long long)(steps * 14 * nthr), w.SU.p, w.mapU.p);
    B200_CUDA(cudaStreamSynchronize(h.stream));
    idx.release();
  }
  w.yin.ensure(steps * nthr); w.y.ensure(steps * nthr); w.x.ensure(steps * nthr);
  B200_CUDA(cudaMemsetAsync(w.yin.p, 0, steps * nthr * sizeof(double), h.stream));
  k_wave_sentinel<<<NUM_SMS * 8, 256, 0, h.stream>>>((long long)(steps * nthr), w.x.p);
  B200_CUDA(cudaGetLastError());
  w.ready = true;
  if (getenv("B200_WAVE_DEBUG"))
    fprintf(stderr, "[wave] grid %d x %d x %d, tiles %d x %d lines, %d strips x %d groups, %d tiles of %d steps, streams %.2f + %.2f GB\n", g.NR, g.NL, g.NP,
            g.TB, g.TC, g.NS, g.NG, g.ntiles, g.NT, steps * 13 * nthr * 8e-9, steps * 14 * nthr * 8e-9);
}

void wave_refresh_values(Handle &h) {
  if (!h.wv.ready || h.n == 0) return;
  const size_t steps = (size_t)h.wv.g.nsteps(), nthr = (size_t)h.wv.g.nthr();
  stream_gather(h, (long long)(steps * 13 * nthr), h.wv.mapL.p, h.d_ilu.p, h.wv.SL.p);
  stream_gather(h, (long long)(steps * 14 * nthr), h.wv.mapU.p, h.d_ilu.p, h.wv.SU.p);
}

template <bool UPPER, int TB, int TC, int E, int PF>
static void wave_launch_cfg(Handle &h, const double *S, const double *rhs, double *out, double *nat) {
  const void *kern = (const void *)k_wave<UPPER, TB, TC, E, PF>;
  constexpr int NTHR = TB * TC, NH = TB + 2 + 2 * TC, NBLK = NTHR + 32 * ((NH + 31) / 32) + 32;
  const size_t smem = (size_t)WV_RING * (TB + 2) * (TC + 1) * 8 + 16;
  int dev = 0, sms = 0, per_sm = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NBLK, smem));
  B200_REQUIRE(per_sm >= 1, "wave-tile triangular solve: kernel does not fit on an SM");
  const int want = h.wv_blocks_per_sm > 0 ? std::min(per_sm, h.wv_blocks_per_sm) : std::min(per_sm, 2);
  const int blocks = std::max(1, std::min(sms * want, h.wv.g.ntiles));
  WaveGeom g = h.wv.g; Ctrl *ctrl = h.ctrl.p; long long *trace = h.wv.trace_on ? h.wv.trace.p : nullptr;
  const int *tile_of = h.wv.tile_of.p, *tsig = h.wv.tile_sig.p, *tgrp = h.wv.tile_grp.p;
  void *argv[] = {(void *)&g, (void *)&tile_of, (void *)&tsig, (void *)&tgrp, (void *)&S, (void *)&rhs, (void *)&out, (void *)&ctrl, (void *)&trace, (void *)&nat};
  B200_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(NBLK), argv, smem, h.stream));
}

template <bool UPPER>
static void wave_launch(Handle &h, const double *S, const double *rhs, double *out, double *nat) {
  switch (h.wv_cfg) {                                                // <tile TB x TC, request lead, prefetch lead>
    case 1: wave_launch_cfg<UPPER, 16, 4, 7, 16>(h, S, rhs, out, nat); break;
    case 2: wave_launch_cfg<UPPER, 16, 8, 7, 0>(h, S, rhs, out, nat); break;
    case 3: wave_launch_cfg<UPPER, 8, 8, 7, 16>(h, S, rhs, out, nat); break;
    case 4: wave_launch_cfg<UPPER, 16, 8, 7, 32>(h, S, rhs, out, nat); break;
    case 5: wave_launch_cfg<UPPER, 32, 4, 7, 16>(h, S, rhs, out, nat); break;
    case 6: wave_launch_cfg<UPPER, 16, 8, 7, 16>(h, S, rhs, out, nat); break;
    default: wave_launch_cfg<UPPER, 16, 8, 3, 16>(h, S, rhs, out, nat); break;     // measured best on 201^3 (request lead 3: 2.12 ms, 7: 2.35 ms)
  }
}

void lu_apply_wave(Handle &h, double *u, const double *v) {
  B200_REQUIRE(h.wv.ready, "wave-tile triangular solve without a plan");
  WavePlan &w = h.wv;
  // Way in: transposing conversion (k_wave_in2, 64 us on the 200^3 problem; B200_WAVE_CONV=1: element-wise k_wave_in, 115 us).  [Reading the
  // right-hand side straight from the natural-order vector inside the forward sweep was measured too: +90 us, its per-thread loads are
  // not covered by the stream prefetch.]  Way out: the backward sweep stores its result in natural order as well and a fill kernel hands
  // the slots back as sentinels (18 us; B200_WAVE_OUT=1: the conversion pass k_wave_out, 210 us).  [Measured and dropped for the way out:
  // a chunk transposition like k_wave_in2 (440 us) and whole lines staged in shared memory and written as one contiguous run (730 us).]
  static const bool direct = !(getenv("B200_WAVE_OUT") && atoi(getenv("B200_WAVE_OUT")) == 1);
  static const bool conv2 = !(getenv("B200_WAVE_CONV") && atoi(getenv("B200_WAVE_CONV")) == 1);
  const int blocks = std::min((h.n + 255) / 256, NUM_SMS * 8);
  const size_t trsm = (size_t)WV_CH * (w.g.nthr() + 1) * sizeof(double);
  const int cblocks = (int)std::min<long long>((long long)w.g.ntiles * ((w.g.NT + WV_CH - 1) / WV_CH), NUM_SMS * 6);
  if (conv2) k_wave_in2<<<cblocks, 256, trsm, h.stream>>>(w.g, w.tile_sig.p, w.tile_grp.p, v, w.yin.p, w.y.p);
  else k_wave_in<<<blocks, 256, 0, h.stream>>>(w.g, w.tile_of.p, h.n, v, w.yin.p, w.g.vlen(), w.y.p);
  wave_launch<false>(h, w.SL.p, w.yin.p, w.y.p, nullptr);
  wave_launch<true>(h, w.SU.p, w.y.p, w.x.p, direct ? u : nullptr);
  if (direct) k_wave_sentinel<<<NUM_SMS * 8, 256, 0, h.stream>>>(w.g.vlen(), w.x.p);
  else k_wave_out<<<blocks, 256, 0, h.stream>>>(w.g, w.tile_of.p, h.n, w.x.p, u);
  B200_CUDA(cudaGetLastError());
  h.st_launch += 4; h.st_pcond++;
}

// per-tile trace (profiles/tools/wave_lab.cu): 8 long long per (sweep, tile): start, end, polls, smid, cycles thread 0 waited for TMA / at the step barrier
void wave_trace_enable(Handle &h, bool on) {
  if (on) {
    const size_t m = (size_t)h.wv.g.ntiles * 2 * 8;
    h.wv.trace.ensure(m);
    B200_CUDA(cudaMemsetAsync(h.wv.trace.p, 0, m * sizeof(long long), h.stream));
  }
  h.wv.trace_on = on;
}
void wave_trace_fetch(Handle &h, std::vector<long long> &out) {
  out.assign((size_t)h.wv.g.ntiles * 2 * 8, 0);
  B200_CUDA(cudaStreamSynchronize(h.stream));
  B200_CUDA(cudaMemcpy(out.data(), h.wv.trace.p, out.size() * sizeof(long long), cudaMemcpyDeviceToHost));
}

}  // namespace b200

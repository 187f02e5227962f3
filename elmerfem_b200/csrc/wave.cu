// Wave-tile triangular solves for the ILU(0) factor of a structured-grid stencil -- CRS_LUSolve, fem/src/CRSMatrix.F90:4590-4663.
// Geometry, data layout and operand routing: wavegeom.h (shared with the CPU emulation tests/wave_harness.cpp).
//
// Why: the level kernel pays one L2 hand-off (0.43 us store -> poll, plus the row itself) per dependency level, 2 x 1401 of them on the
// 200^3 heat problem.  Here a CTA owns a TILE of TB x TC grid lines (one thread per line) that advance in lockstep, one row per thread
// and step: all hand-offs inside a tile go through a shared-memory ring and one CTA barrier per step; only the TB + 2 + 2 TC halo lines
// of a tile come from L2, fetched E steps ahead by loader threads (sentinel protocol: the result vector is pre-filled with a NaN payload
// no arithmetic produces, a loader that still finds it polls).  Shearing the line coordinate (beta = b + c) makes the tile DAG acyclic
// with dependencies only towards smaller (sigma, C), so neighbouring tiles run concurrently, one L2 hop apart, and the critical path
// holds (number of tile rows + columns) hops instead of one per level.  Matrix entries and right-hand sides of a tile step are one
// contiguous block each, moved by TMA bulk copies into an NSLOT-deep shared-memory ring (mbarrier completion).
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

__device__ __forceinline__ unsigned wv_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wv_mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void wv_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wv_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool wv_mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wv_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
__device__ __forceinline__ void wv_st_relaxed(double *p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double wv_ld_relaxed(const double *p) {
  double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ long long wv_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// ---- the sweep ------------------------------------------------------------------------------------------------------------------------
//   forward  (UPPER = false): out_i = rhs_i - sum_{j<i} L_ij out_j                   (4642-4649)
//   backward (UPPER = true) : out_i = Dinv_i * (rhs_i - sum_{j>i} U_ij out_j)       (4653-4660)
// S: matrix stream, RHS: right-hand side at pos(sweep coordinates); Q: result at pos(mirrored sweep coordinates), pre-filled with the
// sentinel.  E: steps a halo row is requested ahead.  Block = TB * TC compute threads + HW loader warps.
template <bool UPPER, int TB, int TC, int NSLOT, int E>
__global__ void __launch_bounds__(TB * TC + 32 * ((TB + 2 + 2 * TC + 31) / 32)) k_wave(WaveGeom g, const int *__restrict__ tile_of, const int *__restrict__ tile_sig,
                                                                                       const int *__restrict__ tile_grp, const double *__restrict__ S,
                                                                                       const double *__restrict__ RHS, double *Q, Ctrl *ctrl, long long *trace) {
  if (ctrl->done) return;
  constexpr int NTHR = TB * TC, NH = TB + 2 + 2 * TC, NE = UPPER ? 14 : 13;
  constexpr int SLOT_D = (NE + 1) * NTHR;                         // doubles per ring slot: matrix rows, then the right-hand sides
  constexpr int YW = TB + 2, YH = TC + 1, YSLOT = YW * YH;
  constexpr int NALL = NTHR + 32 * ((NH + 31) / 32);
  extern __shared__ __align__(128) unsigned char wv_smem[];
  double *ring = reinterpret_cast<double *>(wv_smem);
  double *Yr = ring + (size_t)NSLOT * SLOT_D;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(Yr + WV_RING * YSLOT);
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int q = 0; q < NSLOT; ++q) wv_mbar_init(wv_smem_u32(bars + q), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const int NT = g.NT, NR = g.NR;
  unsigned gs = 0;                                                  // tile steps consumed so far by this CTA (ring position / parity)
  long long spins = 0;
  for (int k = blockIdx.x; k < g.ntiles; k += gridDim.x) {
    const int sig = tile_sig[k], C = tile_grp[k];
    const long long step0 = (long long)k * NT;
    for (int q = tid; q < WV_RING * YSLOT; q += NALL) Yr[q] = 0.0;
    long long tr0 = 0, tr_polls = 0;
    if (trace && tid == NTHR) tr0 = wv_gtime();
    auto issue = [&](int tau, unsigned slot) {                     // step tau of this tile -> ring slot (one thread)
      const unsigned bar = wv_smem_u32(bars + slot);
      wv_mbar_expect_tx(bar, SLOT_D * 8);
      wv_bulk_g2s(wv_smem_u32(ring + (size_t)slot * SLOT_D), S + (step0 + tau) * (NE * NTHR), NE * NTHR * 8, bar);
      wv_bulk_g2s(wv_smem_u32(ring + (size_t)slot * SLOT_D + NE * NTHR), RHS + (step0 + tau) * NTHR, NTHR * 8, bar);
    };
    __syncthreads();                                                // ring zeroed, previous tile's slots all consumed
    if (tid == NTHR) {
#pragma unroll
      for (int q = 0; q < NSLOT; ++q) if (q < NT) issue(q, (gs + q) % NSLOT);
    }
    if (tid < NTHR) {
      // ---------------- compute thread: line (jb, w) ----------------
      const int jb = tid % TB, w = tid / TB;
      const WaveLine ln = wv_line(g, sig, C, jb, w);
      double *qp = Q + (ln.valid ? wv_pos_mirror(g, tile_of, 0, ln.b, ln.c) : 0);     // row a at qp - a * NTHR
      const double *yA = Yr + (w + 1) * YW + (jb + 1), *yB = Yr + w * YW + (jb + 2), *yC = Yr + w * YW + (jb + 1), *yD = Yr + w * YW + jb;
      double *yO = Yr + (w + 1) * YW + (jb + 2);
      double Am = 0.0, A0 = 0.0, Bm = 0.0, B0 = 0.0, Cm = 0.0, C0 = 0.0, Cp = 0.0, Dm = 0.0, D0 = 0.0, Dp = 0.0, h = 0.0;
      double v[NE], rv, acc8 = 0.0, P[8];
      auto fetch = [&](unsigned step) {                            // values of the step into registers, and what does not wait for step - 1
        const unsigned slot = step % NSLOT, bar = wv_smem_u32(bars + slot), parity = (step / NSLOT) & 1u;
        while (!wv_mbar_try_wait(bar, parity)) { if (++spins > WV_SPIN_LIMIT) { ctrl->spin_timeout = 1; break; } }
        const double *sp = ring + (size_t)slot * SLOT_D + tid;
#pragma unroll
        for (int e = 0; e < NE; ++e) v[e] = sp[e * NTHR];
        rv = sp[NE * NTHR];
        if (!UPPER) {
          acc8 = nfms(rv, v[0], Dm); acc8 = nfms(acc8, v[1], D0); acc8 = nfms(acc8, v[2], Dp);
          acc8 = nfms(acc8, v[3], Cm); acc8 = nfms(acc8, v[4], C0); acc8 = nfms(acc8, v[5], Cp);
          acc8 = nfms(acc8, v[6], Bm); acc8 = nfms(acc8, v[7], B0);
        } else {
          P[0] = __dmul_rn(v[0], Dm); P[1] = __dmul_rn(v[1], D0); P[2] = __dmul_rn(v[2], Dp);
          P[3] = __dmul_rn(v[3], Cm); P[4] = __dmul_rn(v[4], C0); P[5] = __dmul_rn(v[5], Cp);
          P[6] = __dmul_rn(v[6], Bm); P[7] = __dmul_rn(v[7], B0);
        }
      };
      fetch(gs);
      for (int tau = 0; tau < NT; ++tau) {
        const double An = yA[((tau - 1) & (WV_RING - 1)) * YSLOT], Bn = yB[((tau - 1) & (WV_RING - 1)) * YSLOT];
        const int a = tau - ln.tau0;
        const bool active = ln.valid && (unsigned)a < (unsigned)NR;
        double acc;
        if (!UPPER) {
          acc = nfms(acc8, v[8], Bn); acc = nfms(acc, v[9], Am); acc = nfms(acc, v[10], A0); acc = nfms(acc, v[11], An); acc = nfms(acc, v[12], h);
        } else {
          acc = nfms(rv, v[12], h); acc = nfms(acc, v[11], An); acc = nfms(acc, v[10], A0); acc = nfms(acc, v[9], Am); acc = nfms(acc, v[8], Bn);
#pragma unroll
          for (int e = 7; e >= 0; --e) acc = __dsub_rn(acc, P[e]);
          acc = __dmul_rn(v[13], acc);
        }
        if (acc != acc) acc = __longlong_as_double((long long)CANON_NAN);
        if (!active) acc = 0.0;
        yO[(tau & (WV_RING - 1)) * YSLOT] = acc;
        if (active) wv_st_relaxed(qp - (long long)a * NTHR, acc);
        h = acc;
        Am = A0; A0 = An; Bm = B0; B0 = Bn;
        // everything of step tau + 1 that does not depend on step tau
        const double Cn = yC[((tau - 2) & (WV_RING - 1)) * YSLOT], Dn = yD[((tau - 4) & (WV_RING - 1)) * YSLOT];
        Cm = C0; C0 = Cp; Cp = Cn; Dm = D0; D0 = Dp; Dp = Dn;
        if (tau + 1 < NT) fetch(gs + tau + 1);
        wv_bar(NALL);
      }
    } else {
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
            wv_bar(NALL);
            // the slot of step tau was read during step tau - 1: free since the previous barrier
            if (tid == NTHR && tau + NSLOT < NT) issue(tau + NSLOT, (gs + tau) % NSLOT);
          }
        }
      }
      if (trace) {
        if (tid == NTHR) { long long *r = trace + ((UPPER ? g.ntiles : 0) + (long long)k) * 4; unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); r[0] = tr0; r[1] = wv_gtime(); r[3] = smid; }
        if (tr_polls) atomicAdd((unsigned long long *)(trace + ((UPPER ? g.ntiles : 0) + (long long)k) * 4 + 2), (unsigned long long)tr_polls);
      }
    }
    gs += NT;
  }
}

// ---- host ------------------------------------------------------------------------------------------------------------------------------
void wave_release(Handle &h) {
  WavePlan &w = h.wv;
  w.SL.release(); w.SU.release(); w.yin.release(); w.y.release(); w.x.release(); w.tile_of.release(); w.tile_sig.release(); w.tile_grp.release(); w.trace.release();
  w.ready = false; w.tried = false;
}

void wave_analyse(Handle &h) {
  WavePlan &w = h.wv;
  if (w.ready || w.tried) return;
  w.tried = true;
  const char *why = nullptr;
  SkewGeom sg;
  if (h.ilu_sep()) why = "ILU(n > 0) / BILU pattern";
  else if (h.nranks > 1) why = "partitioned handle";
  else why = sk_detect(h.n, h.h_rows.data(), h.h_cols.data(), h.h_diag.data(), sg);
  if (why) {
    if (getenv("B200_WAVE_DEBUG")) fprintf(stderr, "[wave] not usable (%s): level kernel stays\n", why);
    return;
  }
  const int cfg = h.wv_cfg;
  const int TB = (cfg == 1) ? 16 : ((cfg == 2) ? 32 : ((cfg == 3) ? 8 : 16)), TC = (cfg == 1) ? 4 : ((cfg == 2) ? 4 : 8);
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
  k_wave_fill<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, h.stream>>>(h.wv.g, h.wv.tile_of.p, h.n, h.d_rows.p, h.d_cols.p, h.d_ilu.p, h.wv.SL.p, h.wv.SU.p);
  B200_CUDA(cudaGetLastError());
}

template <bool UPPER, int TB, int TC, int NSLOT, int E>
static void wave_launch_cfg(Handle &h, const double *S, const double *rhs, double *out) {
  const void *kern = (const void *)k_wave<UPPER, TB, TC, NSLOT, E>;
  constexpr int NTHR = TB * TC, NH = TB + 2 + 2 * TC, NE = UPPER ? 14 : 13, NALL = NTHR + 32 * ((NH + 31) / 32);
  const size_t smem = (size_t)NSLOT * (NE + 1) * NTHR * 8 + (size_t)WV_RING * (TB + 2) * (TC + 1) * 8 + NSLOT * 8;
  int dev = 0, sms = 0, per_sm = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NALL, smem));
  B200_REQUIRE(per_sm >= 1, "wave-tile triangular solve: kernel does not fit on an SM");
  const int want = h.wv_blocks_per_sm > 0 ? std::min(per_sm, h.wv_blocks_per_sm) : per_sm;
  const int blocks = std::max(1, std::min(sms * want, h.wv.g.ntiles));
  WaveGeom g = h.wv.g; Ctrl *ctrl = h.ctrl.p; long long *trace = h.wv.trace_on ? h.wv.trace.p : nullptr;
  const int *tile_of = h.wv.tile_of.p, *tsig = h.wv.tile_sig.p, *tgrp = h.wv.tile_grp.p;
  void *argv[] = {(void *)&g, (void *)&tile_of, (void *)&tsig, (void *)&tgrp, (void *)&S, (void *)&rhs, (void *)&out, (void *)&ctrl, (void *)&trace};
  B200_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(NALL), argv, smem, h.stream));
}

template <bool UPPER>
static void wave_launch(Handle &h, const double *S, const double *rhs, double *out) {
  const WaveGeom &g = h.wv.g;
  if (g.TB == 16 && g.TC == 8) {
    if (h.wv_e == 7) wave_launch_cfg<UPPER, 16, 8, 6, 7>(h, S, rhs, out);
    else if (h.wv_e == 1) wave_launch_cfg<UPPER, 16, 8, 6, 1>(h, S, rhs, out);
    else wave_launch_cfg<UPPER, 16, 8, 6, 3>(h, S, rhs, out);
  }
  else if (g.TB == 16 && g.TC == 4) wave_launch_cfg<UPPER, 16, 4, 8, 3>(h, S, rhs, out);
  else if (g.TB == 32 && g.TC == 4) wave_launch_cfg<UPPER, 32, 4, 6, 3>(h, S, rhs, out);
  else if (g.TB == 8 && g.TC == 8) wave_launch_cfg<UPPER, 8, 8, 8, 3>(h, S, rhs, out);
  else throw Error("wave-tile triangular solve: no kernel for this tile shape");
}

void lu_apply_wave(Handle &h, double *u, const double *v) {
  B200_REQUIRE(h.wv.ready, "wave-tile triangular solve without a plan");
  WavePlan &w = h.wv;
  const int blocks = std::min((h.n + 255) / 256, NUM_SMS * 8);
  k_wave_in<<<blocks, 256, 0, h.stream>>>(w.g, w.tile_of.p, h.n, v, w.yin.p, w.g.vlen(), w.y.p);
  wave_launch<false>(h, w.SL.p, w.yin.p, w.y.p);
  wave_launch<true>(h, w.SU.p, w.y.p, w.x.p);
  k_wave_out<<<blocks, 256, 0, h.stream>>>(w.g, w.tile_of.p, h.n, w.x.p, u);
  B200_CUDA(cudaGetLastError());
  h.st_launch += 4; h.st_pcond++;
}

// per-tile trace (profiles/tools/wave_lab.cu): 4 long long per (sweep, tile): start, end, polls, smid
void wave_trace_enable(Handle &h, bool on) {
  if (on) {
    const size_t m = (size_t)h.wv.g.ntiles * 2 * 4;
    h.wv.trace.ensure(m);
    B200_CUDA(cudaMemsetAsync(h.wv.trace.p, 0, m * sizeof(long long), h.stream));
  }
  h.wv.trace_on = on;
}
void wave_trace_fetch(Handle &h, std::vector<long long> &out) {
  out.assign((size_t)h.wv.g.ntiles * 2 * 4, 0);
  B200_CUDA(cudaStreamSynchronize(h.stream));
  B200_CUDA(cudaMemcpy(out.data(), h.wv.trace.p, out.size() * sizeof(long long), cudaMemcpyDeviceToHost));
}

}  // namespace b200

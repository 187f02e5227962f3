// Jacobi and ILU(0) preconditioning on the device.
//
//   diag_apply   <- CRS_DiagPrecondition          fem/src/CRSMatrix.F90:2279-2326
//   ilu0_factor  <- CRS_IncompleteLU(A,0)         fem/src/CRSMatrix.F90:3445-3531, 3604-3661
//   lu_apply     <- CRS_LUPrecondition/CRS_LUSolve fem/src/CRSMatrix.F90:4550-4564, 4590-4663
//
// The reference runs the factorisation and both triangular sweeps serially.  Here rows are processed
// in dependency-level order by a persistent grid; a row starts as soon as the rows it reads are done
// (point-to-point flags / value sentinels in L2, no grid-wide barrier and no kernel launch per level).
// No reordering, no approximation: every row performs the reference's operations in the reference's
// order with separate multiply/add roundings, so ILUValues and the solve results are bit-identical
// to the CPU loops.
#include "common.cuh"
#include "kernels.cuh"

namespace b200 {

constexpr int ILU_MAXROW = 128;            // rows up to this length are staged in shared memory
constexpr long long SPIN_LIMIT = 1LL << 27;

// ---------------------------------------------------------------------------------------------
__global__ void k_diag_apply(int n, const double *__restrict__ dvals, double *__restrict__ u, const double *__restrict__ v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double d = dvals[i];
    u[i] = (fabs(d) > AEPS) ? __ddiv_rn(v[i], d) : v[i];
  }
}
void diag_apply(Handle &h, double *u, const double *v) {
  if (h.n == 0) return;
  int blocks = std::min((h.n + 255) / 256, NUM_SMS * 8);
  k_diag_apply<<<blocks, 256, 0, h.stream>>>(h.n, h.d_dvals.p, u, v);
  B200_CUDA(cudaGetLastError());
  h.st_launch++;
}

// ---------------------------------------------------------------------------------------------
// One warp per row, rows taken in forward-level order (the L plan's slot order).
__global__ void __launch_bounds__(256) k_ilu0_factor(int nslots, const int *__restrict__ perm, const int *__restrict__ rows,
                                                      const int *__restrict__ cols, const int *__restrict__ diag,
                                                      const double *__restrict__ Avals, double *LU, int *rowdone, Ctrl *ctrl) {
  __shared__ double s_val[8][ILU_MAXROW];
  __shared__ int s_col[8][ILU_MAXROW];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int slot = gwarp; slot < nslots; slot += nwarps) {
    const int r = perm[slot];
    if (r < 0) continue;
    const int rs = rows[r], re = rows[r + 1], d = diag[r], len = re - rs, nlow = d - rs;
    const bool staged = len <= ILU_MAXROW;
    double *vrow = staged ? s_val[wib] : (LU + rs);
    const int *crow = staged ? s_col[wib] : (cols + rs);
    // CRSMatrix.F90:3614-3620: the row in "full form" (here: its own pattern, which is all that is ever touched)
    for (int t = lane; t < len; t += 32) {
      double a = Avals[rs + t];
      if (staged) { s_val[wib][t] = a; s_col[wib][t] = cols[rs + t]; }
      else LU[rs + t] = a;
    }
    __syncwarp();
    // wait until every row of the strict lower pattern is finished
    long long spins = 0;
    for (int t = lane; t < nlow; t += 32) {
      const int k = crow[t];
      while (ld_acquire(rowdone + k) == 0) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        __nanosleep(40);
      }
    }
    __syncwarp();
    // 3624-3637: IKJ elimination, lower entries in column order
    for (int m = 0; m < nlow; ++m) {
      double skm = vrow[m];
      if (skm == 0.0) continue;                                   // 3626
      const int k = crow[m];
      const int kd = diag[k];
      const double ukk = __ldcg(LU + kd);
      if (fabs(ukk) > AEPS) skm = __ddiv_rn(skm, ukk);           // 3628-3629
      __syncwarp();
      if (lane == 0) vrow[m] = skm;
      const int ke = rows[k + 1];
      for (int l = kd + 1 + lane; l < ke; l += 32) {              // 3631-3636
        const int j = cols[l];
        int lo = m + 1, hi = len;                                 // columns > k live right of position m
        while (lo < hi) { int mid = (lo + hi) >> 1; if (crow[mid] < j) lo = mid + 1; else hi = mid; }
        if (lo < len && crow[lo] == j) vrow[lo] = nfms(vrow[lo], skm, __ldcg(LU + l));
      }
      __syncwarp();
    }
    if (staged) for (int t = lane; t < len; t += 32) __stcg(LU + rs + t, s_val[wib][t]);   // 3643-3649
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(rowdone + r, 1);
  }
}

// 3654-3660: store the inverse diagonal (1.0 when tiny)
__global__ void k_ilu0_invert_diag(int n, const int *__restrict__ diag, double *LU) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double d = LU[diag[i]];
    LU[diag[i]] = (fabs(d) < AEPS) ? 1.0 : __ddiv_rn(1.0, d);
  }
}
__global__ void k_gather_diag_slots(int nslots, const int *__restrict__ perm, const int *__restrict__ diag,
                                    const double *__restrict__ LU, double *__restrict__ dinv) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  int r = perm[s];
  dinv[s] = r >= 0 ? LU[diag[r]] : 0.0;
}

static int persistent_blocks(const void *kernel, int threads, int want_per_sm) {
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
  B200_REQUIRE(per_sm > 0, "kernel does not fit on an SM");
  if (want_per_sm > 0 && want_per_sm < per_sm) per_sm = want_per_sm;
  int dev = 0, sms = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return per_sm * sms;
}
// The spin waits need every block of the grid resident at once: cooperative launch guarantees it
// (or fails loudly) even when another stream holds SM resources.
template <class... Args>
static void launch_coresident(const void *kernel, int blocks, int threads, cudaStream_t st, Args... args) {
  void *argv[] = {(void *)&args...};
  B200_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(blocks), dim3(threads), argv, 0, st));
}

void ilu0_factor(Handle &h) {
  B200_REQUIRE(h.have_vals, "ILU0 requested before b200_set_values");
  tri_analyse(h);
  cudaStream_t st = h.stream;
  h.d_ilu.ensure(h.nnz);
  B200_CUDA(cudaEventRecord(h.evf0, st));
  if (h.n > 0) {
    B200_CUDA(cudaMemsetAsync(h.d_rowdone.p, 0, (size_t)h.n * sizeof(int), st));
    B200_CUDA(cudaMemsetAsync(&h.ctrl.p->spin_timeout, 0, sizeof(int), st));
    const double *src = h.have_prec ? h.d_prec.p : h.d_vals.p;        // CRSMatrix.F90:3480-3484
    if (!h.grid_ilu) h.grid_ilu = persistent_blocks((const void *)k_ilu0_factor, 256, 0);
    int blocks = std::max(1, std::min(h.grid_ilu, (h.n_order_f + 7) / 8));
    launch_coresident((const void *)k_ilu0_factor, blocks, 256, st, h.n_order_f, (const int *)h.d_order_f.p, (const int *)h.d_rows.p,
                      (const int *)h.d_cols.p, (const int *)h.d_diag.p, src, h.d_ilu.p, h.d_rowdone.p, h.ctrl.p);
    int eb = std::min((h.n + 255) / 256, NUM_SMS * 8);
    k_ilu0_invert_diag<<<eb, 256, 0, st>>>(h.n, h.d_diag.p, h.d_ilu.p);
    sell_refresh_values(h, h.L, h.d_ilu.p);
    sell_refresh_values(h, h.U, h.d_ilu.p);
    if (h.U.nslots) k_gather_diag_slots<<<(h.U.nslots + 255) / 256, 256, 0, st>>>(h.U.nslots, h.U.perm.p, h.d_diag.p, h.d_ilu.p, h.d_dinv_slot.p);
    B200_CUDA(cudaGetLastError());
  }
  B200_CUDA(cudaEventRecord(h.evf1, st));
  B200_CUDA(cudaMemcpyAsync(h.h_ctrl, h.ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  float ms = 0; B200_CUDA(cudaEventElapsedTime(&ms, h.evf0, h.evf1));
  h.st_factor_ms = ms;
  B200_REQUIRE(h.h_ctrl->spin_timeout == 0, "ILU0 factorisation: dependency wait timed out");
  h.ilu_valid = true; h.ilu_exists = true;
  h.st_factor_launch = h.n > 0 ? 5 : 0;   // factor, invert diag, 2 x SELL refresh, diag gather
}

// ---------------------------------------------------------------------------------------------
// Sync-free level-scheduled triangular solve.  Thread = row, warp = 32 rows of ONE level.  `out` is
// pre-filled with SENTINEL; a consumer spins on the producer's 8-byte result itself, so readiness
// and value arrive in the same L2 round trip.  All gathers of a row are issued together (one L2
// latency per row, not one per entry) and only entries still holding the sentinel are re-polled.
// A single counter of finished slices throttles the grid: a warp starts polling only when every
// slice up to LOOKAHEAD levels behind its own is finished, so the thousands of resident warps that
// run ahead of the wavefront prefetch their matrix entries and then sleep on one address instead of
// flooding L2 with polls.  The counter is only a throttle; correctness rests on the sentinels.
//   forward  (UPPER = false): out_i = rhs_i - sum_{j<i} L_ij out_j                   (4642-4649)
//   backward (UPPER = true) : out_i = Dinv_i * (rhs_i - sum_{j>i} U_ij out_j)       (4653-4660)
__device__ __forceinline__ int ld_relaxed_i(const int *p) {
  int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
// Vectors of the solves live in slot (level) order: `out` has one entry per slot, the column ids of
// T are slot ids.  rhs_idx == nullptr: rhs is in natural order (forward sweep input); otherwise
// rhs[rhs_idx[slot]] (backward sweep reading the forward result).  nat_out != nullptr: the result is
// also scattered to natural order (the preconditioned vector handed back to the Krylov method).
template <bool UPPER, int CH>
__global__ void __launch_bounds__(256, 2) k_sptrsv(SellView T, const int *__restrict__ slice_level, const int *__restrict__ lvl_slices,
                                                    int *lvl_done, int lookahead, unsigned gate_sleep, unsigned spin_sleep,
                                                    const double *__restrict__ dinv_slot, const double *__restrict__ rhs,
                                                    const int *__restrict__ rhs_idx, double *out, double *__restrict__ nat_out,
                                                    Ctrl *ctrl) {
  if (ctrl->done) return;
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int slice = gwarp; slice < T.nslices; slice += nwarps) {
    const long long p0 = T.ptr[slice];
    const int W = (int)((T.ptr[slice + 1] - p0) >> 5);
    const int slot = slice * 32 + lane;
    const int row = T.perm[slot];
    const int len = T.len[slot];
    const int *__restrict__ cp = T.cols + p0 + lane;
    const double *__restrict__ vp = T.vals + p0 + lane;
    double s = row >= 0 ? rhs[rhs_idx ? rhs_idx[slot] : row] : 0.0;
    const double dinv = (UPPER && row >= 0) ? dinv_slot[slot] : 1.0;
    long long spins = 0;
    bool gated = false;
    for (int j0 = 0; j0 < W || !gated; j0 += CH) {
      int c[CH]; double v[CH], xv[CH];
#pragma unroll
      for (int k = 0; k < CH; ++k) {                      // matrix entries: streamed from HBM ahead of the wavefront
        const bool in = j0 + k < W;
        c[k] = in ? ld_stream(cp + (j0 + k) * 32) : 0;
        v[k] = in ? ld_stream(vp + (j0 + k) * 32) : 0.0;
      }
      if (!gated) {                                       // throttle: wait for the wavefront to come near
        const int wl = slice_level[slice] - lookahead;
        if (lane == 0 && wl >= 0) {
          const int need = lvl_slices[wl];
          while (ld_relaxed_i(lvl_done + wl * 32) < need) {
            if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
            if (gate_sleep) __nanosleep(gate_sleep);
          }
        }
        __syncwarp();
        gated = true;
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) xv[k] = (j0 + k < len) ? ld_relaxed(out + c[k]) : 0.0;   // all gathers in flight
      for (;;) {                                          // re-poll every entry still holding the sentinel, together
        bool pending = false;
#pragma unroll
        for (int k = 0; k < CH; ++k) pending |= is_sentinel(xv[k]);
        if (!pending) break;
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        if (spin_sleep) __nanosleep(spin_sleep);
#pragma unroll
        for (int k = 0; k < CH; ++k) if (is_sentinel(xv[k])) xv[k] = ld_relaxed(out + c[k]);
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) if (j0 + k < len) s = nfms(s, v[k], xv[k]);
    }
    if (row >= 0) {
      double res = UPPER ? __dmul_rn(dinv, s) : s;
      if (res != res) res = __longlong_as_double((long long)CANON_NAN);
      st_relaxed(out + slot, res);
      if (nat_out) nat_out[row] = res;
    }
    __syncwarp();
    if (lane == 0) atomicAdd(lvl_done + slice_level[slice] * 32, 1);
  }
}


// ---------------------------------------------------------------------------------------------
// Part-mode triangular solve (default).  One CTA of TRI_NW warps per part (contiguous natural row
// range, see part_layout in structure.cu); warp w of round rho takes slice part_begin + rho*TRI_NW + w:
// 32 rows of ONE dependency level.  Where a row's operands come from is decided at analysis time:
//   * produced by this CTA in the same round     -> shared-memory ring, guarded by an mbarrier token
//   * produced by this CTA up to TRI_RR-2 rounds ago -> shared-memory ring, final since the round barrier
//   * anything else (other CTA, or older)         -> slot-ordered solve vector in L2, "not yet" = SENTINEL
// Same-round hand-off is dataflow: every warp owns one mbarrier; a finishing warp arrives on the
// barriers of the warps that consume its rows (succ mask), a consumer sleeps in mbarrier.try_wait
// (hardware suspend, no issue slots burnt) until all its producers (pred mask) have arrived.  The
// wavefront's critical path therefore runs through shared memory + one mbarrier wake-up per level.
// Operands are consumed strictly left to right with separate multiply/add roundings: bit-identical
// to CRS_LUSolve (CRSMatrix.F90:4642-4660).  Entries before the first same-round operand (npre) are
// accumulated before the warp goes to sleep, so only the tail of the row sum is on the critical path.
// Matrix entries stream from HBM through L2 prefetches issued two rounds ahead.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar, unsigned count) {
  unsigned long long st;
  asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1], %2;" : "=l"(st) : "r"(smem_u32(bar)), "r"(count) : "memory");
  (void)st;
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ double ld_ring(const double *p) {
  double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_u32(p)) : "memory"); return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// code < 0: ~(ring index | same_round << 16)
__device__ __forceinline__ bool code_same_round(int code) { return code < 0 && ((~code) >> 16) != 0; }
__device__ __forceinline__ int code_ring(int code) { return (~code) & 0xffff; }

// in-order row-sum tail s -= p[k] for k = k0 .. 15 (p = v*x already rounded); padded entries carry p = 0
#define B200_TRI_STEP(k) case k: s = __dsub_rn(s, p[k]);
__device__ __forceinline__ double tri_chain_from(double s, const double (&p)[16], int k0) {
  switch (k0) {
    B200_TRI_STEP(0) B200_TRI_STEP(1) B200_TRI_STEP(2) B200_TRI_STEP(3) B200_TRI_STEP(4) B200_TRI_STEP(5) B200_TRI_STEP(6) B200_TRI_STEP(7)
    B200_TRI_STEP(8) B200_TRI_STEP(9) B200_TRI_STEP(10) B200_TRI_STEP(11) B200_TRI_STEP(12) B200_TRI_STEP(13) B200_TRI_STEP(14) B200_TRI_STEP(15)
    default: break;
  }
  return s;
}
#undef B200_TRI_STEP

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ double lds_f64(unsigned addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ int lds_s32(unsigned addr) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }

constexpr int TRI_CH = 16;                                   // entries of an item staged in shared memory
constexpr int TRI_STAGE_BYTES = TRI_CH * 32 * 12;            // values (8 B) then column codes (4 B)
constexpr int TRI_SMEM_BYTES = TRI_RR * TRI_NW * 32 * 8 + 2 * TRI_NW * TRI_STAGE_BYTES + TRI_NW * 8;

// what a warp needs to know about its item besides the matrix entries; fetched one round ahead
struct TriItem { long long p0; int W, len, row, sidx, pred_succ, npre; double b, dinv; };

template <bool UPPER>
__global__ void __launch_bounds__(TRI_NW * 32, 1) k_sptrsv_tok(SellView T, const int2 *__restrict__ meta, const int *__restrict__ part_begin,
                                                                const double *__restrict__ dinv_slot, const double *__restrict__ bslot,
                                                                const int *__restrict__ scatter_idx, double *out, double *scatter_out,
                                                                Ctrl *ctrl, unsigned max_backoff, unsigned long long *trace) {
  extern __shared__ __align__(16) unsigned char tri_smem[];
  double *ring = reinterpret_cast<double *>(tri_smem);
  unsigned char *stages = tri_smem + TRI_RR * TRI_NW * 32 * 8;
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(stages + 2 * TRI_NW * TRI_STAGE_BYTES);
  if (ctrl->done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < TRI_NW) mbar_init(&bar[tid], TRI_NW);
  __syncthreads();
  const int s0 = part_begin[blockIdx.x], s1 = part_begin[blockIdx.x + 1];
  const unsigned ring_base = smem_u32(ring);
  unsigned spins = 0;
  if (trace && tid == 0) trace[(size_t)blockIdx.x * 1024] = gtime();

  auto load_item = [&](int sl) {
    TriItem it;
    const long long a = T.ptr[sl], e = T.ptr[sl + 1];
    const int2 m = meta[sl];
    const int slot = sl * 32 + lane;
    it.p0 = a; it.W = (int)((e - a) >> 5); it.pred_succ = m.x; it.npre = m.y;
    it.len = T.len[slot]; it.row = T.perm[slot]; it.sidx = scatter_idx[slot]; it.b = bslot[slot];
    it.dinv = UPPER ? dinv_slot[slot] : 1.0;
    return it;
  };
  // asynchronous copy of the first TRI_CH entry columns of an item (32 x 8 B values, 32 x 4 B codes each) into a stage
  auto stage_item = [&](const TriItem &it, int st) {
    unsigned char *dst = stages + (size_t)(st * TRI_NW + warp) * TRI_STAGE_BYTES;
    const int w = it.W < TRI_CH ? it.W : TRI_CH;
    const char *vsrc = (const char *)(T.vals + it.p0), *csrc = (const char *)(T.cols + it.p0);
    for (int i = lane; i < w * 16; i += 32) cp_async16(dst + i * 16, vsrc + i * 16);
    for (int i = lane; i < w * 8; i += 32) cp_async16(dst + TRI_CH * 256 + i * 16, csrc + i * 16);
    cp_async_commit();
  };

  TriItem cur, nxt;
  cur.W = 0; nxt.W = 0;
  if (s0 + warp < s1) { cur = load_item(s0 + warp); stage_item(cur, 0); }
  if (s0 + TRI_NW + warp < s1) nxt = load_item(s0 + TRI_NW + warp);

  int rho = 0;
  for (int base = s0; base < s1; base += TRI_NW, ++rho) {
    const int seg = rho % TRI_RR;
    const int slice = base + warp;
    const bool have = slice < s1;
    cp_async_wait_all();
    __syncwarp();
    // next round: entries into the other stage, per-slot data of the round after into registers
    TriItem nn; nn.W = 0;
    if (slice + TRI_NW < s1) stage_item(nxt, (rho + 1) & 1);
    if (slice + 2 * TRI_NW < s1) nn = load_item(slice + 2 * TRI_NW);
    if (have) {
      const unsigned stv = smem_u32(stages + (size_t)((rho & 1) * TRI_NW + warp) * TRI_STAGE_BYTES) + lane * 8;
      const unsigned stc = smem_u32(stages + (size_t)((rho & 1) * TRI_NW + warp) * TRI_STAGE_BYTES) + TRI_CH * 256 + lane * 4;
      const int W = cur.W, len = cur.len, npre = cur.npre;
      const unsigned pred = (unsigned)cur.pred_succ & 0xffffu, succ = (unsigned)cur.pred_succ >> 16;
      double s = cur.b;
      // own token: NW arrivals complete the phase; the popc(pred) producers bring the rest
      if (lane == 0) mbar_arrive(&bar[warp], TRI_NW - __popc(pred));
      bool waited = (pred == 0);
      {
        // ---- staged chunk (entries 0 .. TRI_CH-1): codes from shared memory, operands gathered now unless same-round
        int c[TRI_CH]; double p[TRI_CH];
        unsigned pend = 0, smask = 0;
#pragma unroll
        for (int k = 0; k < TRI_CH; ++k) c[k] = (k < len) ? lds_s32(stc + k * 128) : 0;
#pragma unroll
        for (int k = 0; k < TRI_CH; ++k) {
          double x = 0.0;
          if (k < len) {
            const int code = c[k];
            if (code >= 0) { x = ld_relaxed(out + code); pend |= is_sentinel(x) ? (1u << k) : 0u; }
            else {
              const unsigned addr = ring_base + (unsigned)code_ring(code) * 8u;
              if (waited || !code_same_round(code)) x = lds_f64(addr);
              else { smask |= 1u << k; c[k] = (int)addr; }           // fetched after the token; keep its shared-memory address
            }
          }
          p[k] = x;
        }
        if (__any_sync(0xffffffffu, pend != 0)) {         // another CTA's rows not finished yet: poll L2 with back-off
          unsigned backoff = 32;
          do {
            if (pend) { __nanosleep(backoff); if (backoff < max_backoff) backoff <<= 1; }
#pragma unroll
            for (int k = 0; k < TRI_CH; ++k)
              if ((pend >> k) & 1u) { double x = ld_relaxed(out + c[k]); if (!is_sentinel(x)) { p[k] = x; pend &= ~(1u << k); } }
            if (++spins > (1u << 22)) { ctrl->spin_timeout = 1; pend = 0; }
          } while (__any_sync(0xffffffffu, pend != 0));
        }
#pragma unroll
        for (int k = 0; k < TRI_CH; ++k) p[k] = (k < len && !((smask >> k) & 1u)) ? __dmul_rn(lds_f64(stv + k * 256), p[k]) : 0.0;
        int k0 = 0;
        if (!waited && npre < TRI_CH) {                   // the item's first same-round operand lies in this chunk
          k0 = npre;
#pragma unroll
          for (int k = 0; k < TRI_CH; ++k) { if (k >= k0) break; s = __dsub_rn(s, p[k]); }   // prefix: before going to sleep
          while (!mbar_try_wait(&bar[warp], rho & 1)) { if (++spins > (1u << 22)) { ctrl->spin_timeout = 1; break; } }
          waited = true;
          // ---- critical path of the wavefront from here: same-round operands, tail of the row sum, publish
#pragma unroll
          for (int k = 0; k < TRI_CH; ++k)
            if ((smask >> k) & 1u) p[k] = __dmul_rn(lds_f64(stv + k * 256), lds_f64((unsigned)c[k]));
        }
        s = tri_chain_from(s, p, k0);
      }
      // ---- rows longer than the stage: remaining entries straight from global memory (not the fast path)
      if (W > TRI_CH) {
        const int *__restrict__ cp = T.cols + cur.p0 + lane;
        const double *__restrict__ vp = T.vals + cur.p0 + lane;
        for (int j = TRI_CH; j < W; ++j) {
          const int code = ld_stream(cp + j * 32);
          const double v = ld_stream(vp + j * 32);
          if (!waited && j >= npre) {
            while (!mbar_try_wait(&bar[warp], rho & 1)) { if (++spins > (1u << 22)) { ctrl->spin_timeout = 1; break; } }
            waited = true;
          }
          const bool need = j < len;
          double x = 0.0;
          if (need) x = code >= 0 ? ld_relaxed(out + code) : lds_f64(ring_base + (unsigned)code_ring(code) * 8u);
          unsigned backoff = 32;
          while (__any_sync(0xffffffffu, need && code >= 0 && is_sentinel(x))) {
            if (need && code >= 0 && is_sentinel(x)) {
              if (++spins > (1u << 22)) { ctrl->spin_timeout = 1; x = 0.0; }
              else { __nanosleep(backoff); if (backoff < max_backoff) backoff <<= 1; x = ld_relaxed(out + code); }
            }
          }
          if (need) s = nfms(s, v, x);
        }
      }
      if (!waited) { while (!mbar_try_wait(&bar[warp], rho & 1)) { if (++spins > (1u << 22)) { ctrl->spin_timeout = 1; break; } } }
      double res = UPPER ? __dmul_rn(cur.dinv, s) : s;
      if (res != res) res = __longlong_as_double((long long)CANON_NAN);
      ring[seg * TRI_NW * 32 + tid] = res;                // padded lanes store harmless values nobody reads
      __syncwarp();
      if (lane < TRI_NW && ((succ >> lane) & 1u)) mbar_arrive(&bar[lane], 1);
      // ---- end of the critical path; the copies for other CTAs and for the caller follow
      if (cur.row >= 0) { st_relaxed(out + slice * 32 + lane, res); scatter_out[cur.sidx] = res; }
    }
    __syncthreads();
    cur = nxt; nxt = nn;
    if (trace && tid == 0 && rho < 1022) trace[(size_t)blockIdx.x * 1024 + 1 + rho] = gtime();
  }
  if (trace && tid == 0) trace[(size_t)blockIdx.x * 1024 + 1023] = (unsigned long long)rho;
}

static void lu_launch_part(Handle &h, double *u, const double *v) {
  cudaStream_t st = h.stream;
  unsigned long long *tr = h.d_trace.p;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute((const void *)k_sptrsv_tok<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_SMEM_BYTES));
    B200_CUDA(cudaFuncSetAttribute((const void *)k_sptrsv_tok<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRI_SMEM_BYTES));
    attr_set = true;
  }
  const SellView Lv = h.L.view(), Uv = h.U.view();
  const int2 *mf = (const int2 *)h.d_meta_f.p, *mb = (const int2 *)h.d_meta_b.p;
  const int *pbf = h.d_part_begin_f.p, *pbb = h.d_part_begin_b.p, *l2u = h.d_l2u.p, *up = h.U.perm.p;
  const double *nul = nullptr, *bl = h.d_bl.p, *bu = h.d_bu.p, *dinv = h.d_dinv_slot.p;
  double *yl = h.d_yl.p, *xu = h.d_xu.p, *buw = h.d_bu.p;
  Ctrl *ctrl = h.ctrl.p; unsigned mbk = h.tri_gate_sleep;
  unsigned long long *tr2 = tr ? tr + (size_t)h.tri_parts_f * 1024 : tr;
  { void *argv[] = {(void *)&Lv, (void *)&mf, (void *)&pbf, (void *)&nul, (void *)&bl, (void *)&l2u, (void *)&yl, (void *)&buw, (void *)&ctrl, (void *)&mbk, (void *)&tr};
    B200_CUDA(cudaLaunchCooperativeKernel((const void *)k_sptrsv_tok<false>, dim3(h.tri_parts_f), dim3(TRI_NW * 32), argv, TRI_SMEM_BYTES, st)); }
  { void *argv[] = {(void *)&Uv, (void *)&mb, (void *)&pbb, (void *)&dinv, (void *)&bu, (void *)&up, (void *)&xu, (void *)&u, (void *)&ctrl, (void *)&mbk, (void *)&tr2};
    B200_CUDA(cudaLaunchCooperativeKernel((const void *)k_sptrsv_tok<true>, dim3(h.tri_parts_b), dim3(TRI_NW * 32), argv, TRI_SMEM_BYTES, st)); }
}

// part mode: sentinel fill of the two slot-ordered solve vectors + gather of the right-hand side into L-slot order
__global__ void k_tri_prepare_part(int nsf, double *yl, int nsb, double *xu, const int *__restrict__ permL, const double *__restrict__ v,
                                   double *__restrict__ bl) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, st = gridDim.x * blockDim.x;
  for (int i = i0; i < nsf; i += st) { yl[i] = sent; int r = permL[i]; bl[i] = r >= 0 ? v[r] : 0.0; }
  for (int i = i0; i < nsb; i += st) xu[i] = sent;
}

__global__ void k_tri_prepare(int na, double *a, int nb, double *b, int *counters, int ncounters) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, st = gridDim.x * blockDim.x;
  for (int i = i0; i < na; i += st) a[i] = sent;
  for (int i = i0; i < nb; i += st) b[i] = sent;
  for (int i = i0; i < ncounters; i += st) counters[i * 32] = 0;
}

template <int CH>
static void lu_launch(Handle &h, double *u, const double *v) {
  cudaStream_t st = h.stream;
  if (!h.grid_tri_l) {
    h.grid_tri_l = persistent_blocks((const void *)k_sptrsv<false, CH>, 256, h.tri_blocks_per_sm);
    h.grid_tri_u = persistent_blocks((const void *)k_sptrsv<true, CH>, 256, h.tri_blocks_per_sm);
  }
  int bl = std::max(1, std::min(h.grid_tri_l, (h.L.nslices + 7) / 8));
  int bu = std::max(1, std::min(h.grid_tri_u, (h.U.nslices + 7) / 8));
  const int la = h.tri_lookahead; const unsigned gs = h.tri_gate_sleep, ss = h.tri_spin_sleep;
  launch_coresident((const void *)k_sptrsv<false, CH>, bl, 256, st, h.L.view(), (const int *)h.L.gate.p, (const int *)h.d_lvlcnt_f.p, h.tri_counters.p, la, gs, ss,
                    (const double *)nullptr, v, (const int *)nullptr, h.d_yl.p, (double *)nullptr, h.ctrl.p);
  launch_coresident((const void *)k_sptrsv<true, CH>, bu, 256, st, h.U.view(), (const int *)h.U.gate.p, (const int *)h.d_lvlcnt_b.p,
                    h.tri_counters.p + (size_t)(h.nlev_f + 1) * 32, la, gs, ss,
                    (const double *)h.d_dinv_slot.p, (const double *)h.d_yl.p, (const int *)h.d_urhs.p, h.d_xu.p, u, h.ctrl.p);
}

void lu_apply(Handle &h, double *u, const double *v) {
  B200_REQUIRE(h.ilu_valid, "LU preconditioner applied without a valid ILU0 factor");
  if (h.n == 0) return;
  if (h.tri_mode == 1 && getenv("B200_TRI_TRACE") && !h.d_trace.p) {
    h.d_trace.ensure((size_t)(h.tri_parts_f + h.tri_parts_b) * 1024);
    B200_CUDA(cudaMemset(h.d_trace.p, 0, (size_t)(h.tri_parts_f + h.tri_parts_b) * 1024 * sizeof(unsigned long long)));
  }
  if (h.tri_mode == 1) {
    k_tri_prepare_part<<<NUM_SMS * 8, 256, 0, h.stream>>>(h.L.nslots, h.d_yl.p, h.U.nslots, h.d_xu.p, h.L.perm.p, v, h.d_bl.p);
    lu_launch_part(h, u, v);
  } else {
    k_tri_prepare<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, h.stream>>>(h.L.nslots, h.d_yl.p, h.U.nslots, h.d_xu.p, h.tri_counters.p, h.nlev_f + h.nlev_b + 2);
    if (h.tri_maxw <= 8) lu_launch<8>(h, u, v); else lu_launch<16>(h, u, v);
  }
  B200_CUDA(cudaGetLastError());
  h.st_launch += 3; h.st_pcond++;
  if (h.d_trace.p && h.tri_mode == 1) {                 // B200_TRI_TRACE=<file>: per-CTA round time stamps of the last application
    const char *path = getenv("B200_TRI_TRACE");
    size_t cnt = (size_t)(h.tri_parts_f + h.tri_parts_b) * 1024;
    std::vector<unsigned long long> t(cnt);
    B200_CUDA(cudaStreamSynchronize(h.stream));
    B200_CUDA(cudaMemcpy(t.data(), h.d_trace.p, cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (FILE *f = fopen(path, "wb")) { int hdr[2] = {h.tri_parts_f, h.tri_parts_b}; fwrite(hdr, sizeof hdr, 1, f); fwrite(t.data(), sizeof(unsigned long long), cnt, f); fclose(f); }
  }
}

}  // namespace b200

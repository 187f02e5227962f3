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
__device__ __forceinline__ int ld_relaxed_i(const int *p) {
  int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
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
                                                      const double *__restrict__ Avals, const int *__restrict__ src, double *LU, int *rowdone,
                                                      Ctrl *ctrl) {
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
      double a;                                                   // ILU(n > 0): src maps the pattern to the matrix, -1 = fill
      if (src) { const int q = src[rs + t]; a = q >= 0 ? Avals[q] : 0.0; } else a = Avals[rs + t];
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
    // 3624-3637: IKJ elimination, lower entries in column order.  The pivot rows are complete, so everything the
    // elimination reads from them (the pivots u_kk and the upper parts of the rows k) is fetched for a
    // batch of MB lower entries at once -- one L2 round trip per batch instead of two dependent ones per entry
    // (35 ms -> the factorisation of C2 was 13 x 2 round trips per row) -- and the arithmetic then runs from registers
    // in the reference's order.
    constexpr int MB = 8;
    for (int m0 = 0; m0 < nlow; m0 += MB) {
      int mkd = 0, mke = 0; double mukk = 0.0;
      if (lane < MB && m0 + lane < nlow) { const int k = crow[m0 + lane]; mkd = diag[k]; mke = rows[k + 1]; mukk = __ldcg(LU + mkd); }
      int pj0[MB], pj1[MB]; double pv0[MB], pv1[MB];
#pragma unroll
      for (int t = 0; t < MB; ++t) {
        const int kd = __shfl_sync(0xffffffffu, mkd, t), ke = __shfl_sync(0xffffffffu, mke, t);
        const int l0 = kd + 1 + lane, l1 = l0 + 32;
        const bool in0 = (m0 + t < nlow) && l0 < ke, in1 = (m0 + t < nlow) && l1 < ke;
        pj0[t] = in0 ? cols[l0] : -1; pv0[t] = in0 ? __ldcg(LU + l0) : 0.0;
        pj1[t] = in1 ? cols[l1] : -1; pv1[t] = in1 ? __ldcg(LU + l1) : 0.0;
      }
#pragma unroll
      for (int t = 0; t < MB; ++t) {
        const int m = m0 + t;
        if (m >= nlow) break;
        double skm = vrow[m];
        const double ukk = __shfl_sync(0xffffffffu, mukk, t);
        const int kd = __shfl_sync(0xffffffffu, mkd, t), ke = __shfl_sync(0xffffffffu, mke, t);
        if (skm == 0.0) continue;                                   // 3626
        if (fabs(ukk) > AEPS) skm = __ddiv_rn(skm, ukk);           // 3628-3629
        __syncwarp();
        if (lane == 0) vrow[m] = skm;
        auto update = [&](int j, double ukj) {                      // 3631-3636
          int lo = m + 1, hi = len;                                 // columns > k live right of position m
          while (lo < hi) { int mid = (lo + hi) >> 1; if (crow[mid] < j) lo = mid + 1; else hi = mid; }
          if (lo < len && crow[lo] == j) vrow[lo] = nfms(vrow[lo], skm, ukj);
        };
        if (pj0[t] >= 0) update(pj0[t], pv0[t]);
        if (pj1[t] >= 0) update(pj1[t], pv1[t]);
        for (int l = kd + 1 + 64 + lane; l < ke; l += 32) update(cols[l], __ldcg(LU + l));   // pivot rows wider than 64 upper entries
        __syncwarp();
      }
    }
    if (staged) for (int t = lane; t < len; t += 32) __stcg(LU + rs + t, s_val[wib][t]);   // 3643-3649
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(rowdone + r, 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Incomplete Cholesky, CRS_IncompleteLU with A % Cholesky (CRSMatrix.F90:3539-3602; 'Linear System Symmetric ILU').  One warp per row in
// forward-level order, the lower part of the row (T, then S) staged in shared memory.  For the m-th lower entry j, in column order:
//   S(j) = (T(j) - sum_l S(k_l) L(j, k_l)) * Linv(j, j)  over the whole lower part of row j, S(k) = 0 outside row i's pattern,
//   S(i) = S(i) - S(j)^2;                        finally Linv(i, i) = 1 / sqrt(S(i))   (1 when S(i) <= AEPS).
// The lanes form the products S(k_l) L(j, k_l) of 32 entries of row j at once (one rounding each, as the reference); the subtractions
// run in the reference's order from shuffled products.  Only the lower part and the diagonal are written (upper part: 0).
__global__ void __launch_bounds__(256) k_ichol_factor(int nslots, const int *__restrict__ perm, const int *__restrict__ rows, const int *__restrict__ cols,
                                                       const int *__restrict__ diag, const double *__restrict__ Avals, const int *__restrict__ src, double *LU,
                                                       int *rowdone, Ctrl *ctrl) {
  __shared__ double s_val[8][ILU_MAXROW];
  __shared__ int s_col[8][ILU_MAXROW];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int slot = gwarp; slot < nslots; slot += nwarps) {
    const int r = perm[slot];
    if (r < 0) continue;
    const int rs = rows[r], re = rows[r + 1], d = diag[r], nlow = d - rs;       // nlow <= ILU_MAXROW - 1 (checked by the host)
    for (int t = lane; t <= nlow; t += 32) {                                    // 3553-3556, 3566: T, on the factor's pattern
      double a;
      if (src) { const int q = src[rs + t]; a = q >= 0 ? Avals[q] : 0.0; } else a = Avals[rs + t];
      s_val[wib][t] = a; s_col[wib][t] = cols[rs + t];
    }
    __syncwarp();
    long long spins = 0;
    for (int t = lane; t < nlow; t += 32) {
      const int k = s_col[wib][t];
      while (ld_acquire(rowdone + k) == 0) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        __nanosleep(40);
      }
    }
    __syncwarp();
    double Si = s_val[wib][nlow];
    for (int m = 0; m < nlow; ++m) {                                            // 3567-3576
      const int j = s_col[wib][m], jr = rows[j], jd = diag[j], jl = jd - jr;
      double acc = s_val[wib][m];
      for (int c0 = 0; c0 < jl; c0 += 32) {
        const int l = c0 + lane;
        double prod = 0.0;
        if (l < jl) {
          const int k = cols[jr + l];
          const double Lv = __ldcg(LU + jr + l);
          int lo = 0, hi = m;                                                   // columns < j of row i sit left of position m
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_col[wib][mid] < k) lo = mid + 1; else hi = mid; }
          const double Sk = (lo < m && s_col[wib][lo] == k) ? s_val[wib][lo] : 0.0;
          prod = __dmul_rn(Sk, Lv);
        }
        const int cnt = jl - c0 < 32 ? jl - c0 : 32;
        for (int t = 0; t < cnt; ++t) acc = __dsub_rn(acc, __shfl_sync(0xffffffffu, prod, t));
      }
      acc = __dmul_rn(acc, __ldcg(LU + jd));
      __syncwarp();
      if (lane == 0) s_val[wib][m] = acc;
      __syncwarp();
      Si = __dsub_rn(Si, __dmul_rn(acc, acc));
    }
    if (Si <= AEPS) Si = 1.0;                                                   // 3578-3587
    else Si = __ddiv_rn(1.0, __dsqrt_rn(Si));
    for (int t = lane; t < nlow; t += 32) __stcg(LU + rs + t, s_val[wib][t]);   // 3596-3601
    if (lane == 0) __stcg(LU + d, Si);
    for (int t = d + 1 + lane; t < re; t += 32) __stcg(LU + t, 0.0);
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(rowdone + r, 1);
  }
}


// ---- position map: the symbolic half of the elimination, once per structure -----------------------------------------------------------
// For row i, its m-th lower entry (pivot row k) and the u-th upper entry (column j) of row k: the position of column j in row i, or 255.
// k_ilu0_factor finds that position with a binary search over the staged row on every update and every factorisation (85 instructions
// per stored entry, profiles/r01_ncu_summary.txt); with the map a refactorisation only streams one byte per update.  Layout: row i owns
// nlow_i * maxu bytes at posptr[i].  Built when it fits (see ilu0_factor); otherwise the searching kernel stays.
__global__ void __launch_bounds__(256) k_ilu_posmap(int n, const int *__restrict__ rows, const int *__restrict__ cols, const int *__restrict__ diag,
                                                     const long long *__restrict__ posptr, int maxu, unsigned char *__restrict__ pos) {
  const int lane = threadIdx.x & 31;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += (gridDim.x * blockDim.x) >> 5) {
    const int rs = rows[i], len = rows[i + 1] - rs, nlow = diag[i] - rs;
    unsigned char *out = pos + posptr[i];
    for (int m = 0; m < nlow; ++m) {
      const int k = cols[rs + m], kd = diag[k], ke = rows[k + 1];
      for (int u = lane; u < maxu; u += 32) {
        unsigned char p = 255;
        if (kd + 1 + u < ke) {
          const int j = cols[kd + 1 + u];
          int lo = m + 1, hi = len;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (cols[rs + mid] < j) lo = mid + 1; else hi = mid; }
          if (lo < len && cols[rs + lo] == j) p = (unsigned char)lo;
        }
        out[(size_t)m * maxu + u] = p;
      }
    }
  }
}
// The factorisation with the map (rows staged in shared memory, pivot rows of <= 64 upper entries): same operations on the same operands
// in the same order as k_ilu0_factor -- bit-identical ILUValues.
__global__ void __launch_bounds__(256) k_ilu0_factor_map(int nslots, const int *__restrict__ perm, const int *__restrict__ rows,
                                                          const int *__restrict__ cols, const int *__restrict__ diag,
                                                          const double *__restrict__ Avals, const int *__restrict__ src, double *LU, int *rowdone,
                                                          Ctrl *ctrl, const long long *__restrict__ posptr, int maxu, const unsigned char *__restrict__ pos) {
  __shared__ double s_val[8][ILU_MAXROW];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double *vrow = s_val[wib];
  for (int slot = gwarp; slot < nslots; slot += nwarps) {
    const int r = perm[slot];
    if (r < 0) continue;
    const int rs = rows[r], re = rows[r + 1], d = diag[r], len = re - rs, nlow = d - rs;
    const unsigned char *pm = pos + posptr[r];
    for (int t = lane; t < len; t += 32) {
      double a;
      if (src) { const int q = src[rs + t]; a = q >= 0 ? Avals[q] : 0.0; } else a = Avals[rs + t];
      vrow[t] = a;
    }
    long long spins = 0;
    for (int t = lane; t < nlow; t += 32) {
      const int k = cols[rs + t];
      while (ld_acquire(rowdone + k) == 0) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        __nanosleep(40);
      }
    }
    __syncwarp();
    constexpr int MB = 8;
    for (int m0 = 0; m0 < nlow; m0 += MB) {
      int mkd = 0, mke = 0; double mukk = 0.0;
      if (lane < MB && m0 + lane < nlow) { const int k = cols[rs + m0 + lane]; mkd = diag[k]; mke = rows[k + 1]; mukk = __ldcg(LU + mkd); }
      unsigned char pp0[MB], pp1[MB]; double pv0[MB], pv1[MB];
#pragma unroll
      for (int t = 0; t < MB; ++t) {
        const int kd = __shfl_sync(0xffffffffu, mkd, t), ke = __shfl_sync(0xffffffffu, mke, t);
        const int l0 = kd + 1 + lane, l1 = l0 + 32;
        const bool in0 = (m0 + t < nlow) && l0 < ke, in1 = (m0 + t < nlow) && l1 < ke;
        pp0[t] = in0 ? pm[(size_t)(m0 + t) * maxu + lane] : (unsigned char)255; pv0[t] = in0 ? __ldcg(LU + l0) : 0.0;
        pp1[t] = in1 ? pm[(size_t)(m0 + t) * maxu + 32 + lane] : (unsigned char)255; pv1[t] = in1 ? __ldcg(LU + l1) : 0.0;
      }
#pragma unroll
      for (int t = 0; t < MB; ++t) {
        const int m = m0 + t;
        if (m >= nlow) break;
        double skm = vrow[m];
        const double ukk = __shfl_sync(0xffffffffu, mukk, t);
        if (skm == 0.0) continue;                                   // 3626
        if (fabs(ukk) > AEPS) skm = __ddiv_rn(skm, ukk);           // 3628-3629
        __syncwarp();
        if (lane == 0) vrow[m] = skm;
        if (pp0[t] != 255) vrow[pp0[t]] = nfms(vrow[pp0[t]], skm, pv0[t]);   // 3631-3636: every target column once per pivot row
        if (pp1[t] != 255) vrow[pp1[t]] = nfms(vrow[pp1[t]], skm, pv1[t]);
        __syncwarp();
      }
    }
    for (int t = lane; t < len; t += 32) __stcg(LU + rs + t, vrow[t]);   // 3643-3649
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
  B200_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(blocks), dim3(threads), argv, 0, st));   // (a plain launch measured no faster)
}

static void tri_autotune_wave(Handle &h);

// The factorisation for narrow rows (<= 32 entries, <= 16 of them left of the diagonal, pivot rows of <= 32 upper entries: scalar 27-point
// stencils) with the row in REGISTERS, one entry per lane.  What limits k_ilu0_factor_map is not arithmetic but the chain of dependent
// memory round trips a row makes AFTER its pivot rows are finished (1401 levels x 23 us on the 200^3 problem).  Here everything that does
// not depend on the pivot rows' values -- the row itself, the pivot rows' extents, the position map, turned from scatter form (entry u of
// pivot row m -> position in this row) into gather form (position -> entry u) through shared memory -- is fetched BEFORE the wait on the
// rowdone flags; after it a single round trip brings the pivots and, per lane, the one operand of each pivot row that lands on the lane's
// position; the elimination then runs from registers with one shuffle + one division per pivot (no shared memory, no warp barrier), and the
// row is published.  Same operations on the same operands in the same order as k_ilu0_factor: bit-identical ILUValues.
__global__ void __launch_bounds__(256, 4) k_ilu0_factor_reg(int nslots, const int *__restrict__ perm, const int *__restrict__ rows, const int *__restrict__ cols,
                                                          const int *__restrict__ diag, const double *__restrict__ Avals, const int *__restrict__ src, double *LU,
                                                          int *rowdone, Ctrl *ctrl, const long long *__restrict__ posptr, int maxu,
                                                          const unsigned char *__restrict__ pos) {
  constexpr int NM = 16;
  __shared__ unsigned char s_inv[8][NM][32];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int slot = gwarp; slot < nslots; slot += nwarps) {
    const int r = perm[slot];
    if (r < 0) continue;
    const int rs = rows[r], len = rows[r + 1] - rs, nlow = diag[r] - rs;
    const unsigned char *pm = pos + posptr[r];
    // ---- before the wait: the row, the pivot rows' diagonals' positions, the gather form of the position map
    double v = 0.0;
    if (lane < len) { if (src) { const int q = src[rs + lane]; v = q >= 0 ? Avals[q] : 0.0; } else v = Avals[rs + lane]; }
    const int kcol = lane < nlow ? cols[rs + lane] : -1;
    const int kd = kcol >= 0 ? diag[kcol] : 0;
    __syncwarp();
#pragma unroll
    for (int m = 0; m < NM; ++m) s_inv[wib][m][lane] = 255;
    __syncwarp();
#pragma unroll
    for (int m = 0; m < NM; ++m)
      if (m < nlow && lane < maxu) { const unsigned char pp = pm[(size_t)m * maxu + lane]; if (pp != 255) s_inv[wib][m][pp] = (unsigned char)lane; }
    __syncwarp();
    unsigned char gi[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) gi[m] = s_inv[wib][m][lane];
    // ---- wait until every row of the strict lower pattern is finished
    long long spins = 0;
    if (kcol >= 0) {
      while (ld_acquire(rowdone + kcol) == 0) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        __nanosleep(20);
      }
    }
    __syncwarp();
    // ---- one round trip: pivots, and per lane the operand of each pivot row that meets the lane's position
    const double ukk_l = kcol >= 0 ? __ldcg(LU + kd) : 0.0;
    double pv[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const int kdm = __shfl_sync(FULL, kd, m);
      pv[m] = (m < nlow && gi[m] != 255) ? __ldcg(LU + kdm + 1 + gi[m]) : 0.0;
    }
    // ---- 3624-3637 from registers
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      if (m < nlow) {
        double skm = __shfl_sync(FULL, v, m);
        const double ukk = __shfl_sync(FULL, ukk_l, m);
        if (skm != 0.0) {                                           // 3626
          if (fabs(ukk) > AEPS) skm = __ddiv_rn(skm, ukk);         // 3628-3629
          if (lane == m) v = skm;
          if (gi[m] != 255) v = nfms(v, skm, pv[m]);               // 3631-3636
        }
      }
    }
    if (lane < len) __stcg(LU + rs + lane, v);                      // 3643-3649
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(rowdone + r, 1);
  }
}

// CRS_IncompleteLU with A % Cholesky set: the factor lives in the lower part + diagonal of d_ilu (same pattern as the LU factor)
static void ichol_factor(Handle &h) {
  ichol_analyse(h);
  cudaStream_t st = h.stream;
  h.d_ilu.ensure(h.lnnz());
  B200_CUDA(cudaEventRecord(h.evf0, st));
  if (h.n > 0) {
    const std::vector<int> &R = h.lrows(), &Dg = h.ldiag();
    int maxlow = 0;
    for (int i = 0; i < h.n; ++i) maxlow = std::max(maxlow, Dg[i] - R[i]);
    B200_REQUIRE(maxlow + 1 <= ILU_MAXROW, "incomplete Cholesky: a row has more than 127 entries left of the diagonal");
    B200_CUDA(cudaMemsetAsync(h.d_rowdone.p, 0, (size_t)h.n * sizeof(int), st));
    B200_CUDA(cudaMemsetAsync(&h.ctrl.p->spin_timeout, 0, sizeof(int), st));
    const double *src = h.have_prec ? h.d_prec.p : h.d_vals.p;        // CRSMatrix.F90:3480-3484
    static int grid = 0;
    if (!grid) grid = persistent_blocks((const void *)k_ichol_factor, 256, 0);
    const int blocks = std::max(1, std::min(grid, (h.L.nslots + 7) / 8));
    launch_coresident((const void *)k_ichol_factor, blocks, 256, st, h.L.nslots, (const int *)h.L.perm.p, h.d_lrows(), h.d_lcols(), h.d_ldiag(), src,
                      (const int *)(h.ilu_sep() ? h.dl_src.p : nullptr), h.d_ilu.p, h.d_rowdone.p, h.ctrl.p);
    B200_CUDA(cudaGetLastError());
  }
  B200_CUDA(cudaEventRecord(h.evf1, st));
  B200_CUDA(cudaMemcpyAsync(h.h_ctrl, h.ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  float ms = 0; B200_CUDA(cudaEventElapsedTime(&ms, h.evf0, h.evf1));
  h.st_factor_ms = ms;
  B200_REQUIRE(h.h_ctrl->spin_timeout == 0, "incomplete Cholesky factorisation: dependency wait timed out");
  h.ilu_valid = true; h.ilu_exists = true;
  h.st_factor_launch = h.n > 0 ? 1 : 0;
  h.tri_mode = 0;                                                   // the Cholesky sweeps are the only solve path of this factor
}

// ---------------------------------------------------------------------------------------------
// ILUT, CRS_ILUT / ComputeILUT (CRSMatrix.F90:4144-4340; 'Linear System Preconditioning = ILUT', 'Linear System ILUT Tolerance').  The
// pattern of the factor is decided by the values, row by row, so there is no symbolic phase and no level schedule: one warp per row in
// NATURAL order (warp w takes rows w, w + W, ...), a row waits for each of its pivot rows on its rowdone flag as it reaches it -- every
// dependency points to a smaller row and every warp walks upwards, so the smallest unfinished row never waits and the wavefront forms by
// itself.  The working row lives in shared memory as a column-sorted list (the reference's flagged full-length vector, 4231-4258): pivots
// are taken in ascending column order; the upper part of a finished pivot row is merged in 32 entries at a time -- present columns are
// updated in place (S(j) = S(j) - S(k) U(k,j)), absent ones are inserted at their sorted position with S(j) = 0 - S(k) U(k,j), which is what
// the reference's zero-initialised vector gives.  Then the drop rule (keep |S| >= TOL * ||A(i,:)||_2, always the diagonal; the norm summed
// in storage order) and the row goes to a pool at an atomically reserved offset.  Same operations, same order: bit-identical factors.
constexpr int ILUT_CAP = 1024;        // entries of a working row
constexpr int ILUT_WARPS = 4;
__global__ void __launch_bounds__(ILUT_WARPS * 32) k_ilut_factor(int n, const int *__restrict__ rows, const int *__restrict__ cols, const double *__restrict__ Avals,
                                                                  double tol, int *pcols, double *pvals, unsigned long long cap, unsigned long long *top,
                                                                  long long *rstart, int *rlen, int *rdiag, int *rowdone, Ctrl *ctrl, int *overflow) {
  extern __shared__ __align__(16) unsigned char ilut_sm[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int *scol = reinterpret_cast<int *>(ilut_sm) + wib * ILUT_CAP;
  double *sval = reinterpret_cast<double *>(ilut_sm + (size_t)ILUT_WARPS * ILUT_CAP * sizeof(int)) + wib * ILUT_CAP;
  int *nip = reinterpret_cast<int *>(ilut_sm + (size_t)ILUT_WARPS * ILUT_CAP * (sizeof(int) + sizeof(double))) + wib * 32;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  long long spins = 0;
  for (int i = gwarp; i < n; i += nwarps) {
    const int rs = rows[i], len = rows[i + 1] - rs;
    bool dead = ld_relaxed_i(overflow) != 0;                        // a row or the pool overflowed somewhere: publish and leave
    if (!dead && len > ILUT_CAP) { if (lane == 0) atomicExch(overflow, 2); dead = true; }
    int L = dead ? 0 : len;
    for (int t = lane; t < L; t += 32) { scol[t] = cols[rs + t]; sval[t] = Avals[rs + t]; }
    __syncwarp();
    double s2 = 0.0;                                                // 4262, before the row is touched
    for (int t = 0; t < L; ++t) { const double a = fabs(sval[t]); s2 = __dadd_rn(s2, __dmul_rn(a, a)); }
    const double thr = __dmul_rn(tol, __dsqrt_rn(s2));
    int pos = 0;
    while (!dead && pos < L) {
      const int k = scol[pos];
      if (k >= i) break;
      while (ld_acquire(rowdone + k) == 0) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        __nanosleep(40);
      }
      if (ld_relaxed_i(overflow) != 0) { dead = true; break; }
      const long long ks = rstart[k];
      const int kl = rlen[k], kdg = rdiag[k];
      const double piv = __ldcg(pvals + ks + kdg);
      double Sk = sval[pos];
      if (fabs(piv) > AEPS) Sk = __ddiv_rn(Sk, piv);               // 4244-4245
      __syncwarp();
      if (lane == 0) sval[pos] = Sk;
      __syncwarp();
      const int U = kl - kdg - 1;
      const long long base = ks + kdg + 1;
      for (int c0 = 0; c0 < U && !dead; c0 += 32) {                 // 4247-4254
        const int e = c0 + lane;
        const bool have = e < U;
        const int j = have ? __ldcg(pcols + base + e) : 0x7fffffff;
        const double v = have ? __ldcg(pvals + base + e) : 0.0;
        int lo = pos + 1, hi = L;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (scol[mid] < j) lo = mid + 1; else hi = mid; }
        const bool found = have && lo < L && scol[lo] == j;
        if (found) sval[lo] = nfms(sval[lo], Sk, v);
        const unsigned nf = __ballot_sync(FULL, have && !found);
        if (nf) {
          const int cnt = __popc(nf), r = __popc(nf & ((1u << lane) - 1u));
          if (L + cnt > ILUT_CAP) { if (lane == 0) atomicExch(overflow, 2); dead = true; break; }
          if (have && !found) nip[r] = lo;
          __syncwarp();
          for (int hiq = L; hiq > pos + 1; hiq -= 32) {             // make room: tail elements move up by the number of new columns below them
            const int q = hiq - 1 - lane;
            const bool in = q >= pos + 1;
            int cq = 0, sh = 0; double vq = 0.0;
            if (in) {
              cq = scol[q]; vq = sval[q];
              int a = 0, b = cnt;
              while (a < b) { const int m = (a + b) >> 1; if (nip[m] <= q) a = m + 1; else b = m; }
              sh = a;
            }
            __syncwarp();
            if (in && sh) { scol[q + sh] = cq; sval[q + sh] = vq; }
            __syncwarp();
          }
          if (have && !found) { scol[lo + r] = j; sval[lo + r] = nfms(0.0, Sk, v); }
          L += cnt;
        }
        __syncwarp();
      }
      ++pos;
    }
    // ---- 4264-4277: drop rule, the row goes to the pool in column order
    int K = 0;
    if (!dead)
      for (int c0 = 0; c0 < L; c0 += 32) {
        const int t = c0 + lane;
        const bool keep = t < L && (fabs(sval[t]) >= thr || scol[t] == i);
        K += __popc(__ballot_sync(FULL, keep));
      }
    unsigned long long off = 0;
    if (lane == 0 && K) off = atomicAdd(top, (unsigned long long)K);
    off = __shfl_sync(FULL, off, 0);
    if (!dead && off + (unsigned long long)K > cap) { if (lane == 0) atomicCAS(overflow, 0, 1); dead = true; }
    int dg = -1;
    if (!dead) {
      int run = 0;
      for (int c0 = 0; c0 < L; c0 += 32) {
        const int t = c0 + lane;
        const bool keep = t < L && (fabs(sval[t]) >= thr || scol[t] == i);
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) {
          const int w = run + __popc(m & ((1u << lane) - 1u));
          __stcg(pcols + off + w, scol[t]); __stcg(pvals + off + w, sval[t]);
          if (scol[t] == i) dg = w;
        }
        run += __popc(m);
      }
      dg = __reduce_max_sync(FULL, dg);
    }
    if (lane == 0) { rstart[i] = (long long)off; rlen[i] = dead ? 0 : K; rdiag[i] = dg; }
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(rowdone + i, 1);
  }
}
// pool -> CRS order on the factor's own pattern
__global__ void k_ilut_gather(int n, const int *__restrict__ orows, const long long *__restrict__ rstart, const int *__restrict__ rdiag, const int *__restrict__ pcols,
                              const double *__restrict__ pvals, int *__restrict__ ocols, double *__restrict__ ovals, int *__restrict__ odiag) {
  const int lane = threadIdx.x & 31;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += (gridDim.x * blockDim.x) >> 5) {
    const int o = orows[i], len = orows[i + 1] - o;
    const long long s = rstart[i];
    for (int t = lane; t < len; t += 32) { ocols[o + t] = pcols[s + t]; ovals[o + t] = pvals[s + t]; }
    if (lane == 0) odiag[i] = o + rdiag[i];
  }
}

static void ilut_factor(Handle &h) {
  cudaStream_t st = h.stream;
  const int n = h.n;
  B200_REQUIRE(h.nranks >= 1, "ILUT: bad handle");
  B200_CUDA(cudaEventRecord(h.evf0, st));
  h.ilu_pat_ready = false; h.tri_ready = false;
  wave_release(h); lane_release(h); ichol_release(h);
  if (n > 0) {
    const double *src = h.have_prec ? h.d_prec.p : h.d_vals.p;        // CRSMatrix.F90:4203-4207
    DBuf<int> pcols, rlen, rdiag, ovf; DBuf<double> pvals; DBuf<long long> rstart; DBuf<unsigned long long> top;
    rlen.ensure(n); rdiag.ensure(n); rstart.ensure(n); ovf.ensure(1); top.ensure(1);
    h.d_rowdone.ensure(n);
    const size_t smem = (size_t)ILUT_WARPS * (ILUT_CAP * (sizeof(int) + sizeof(double)) + 32 * sizeof(int));
    B200_CUDA(cudaFuncSetAttribute((const void *)k_ilut_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, dev = 0, sms = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)k_ilut_factor, ILUT_WARPS * 32, smem));
    B200_REQUIRE(per_sm > 0, "ILUT: kernel does not fit on an SM");
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = std::max(1, std::min(per_sm * sms, (n + ILUT_WARPS - 1) / ILUT_WARPS));
    unsigned long long cap = (unsigned long long)std::max<long long>(4 * h.nnz, 4096);
    int ov = 0;
    for (int attempt = 0; attempt < 6; ++attempt) {
      pcols.ensure(cap); pvals.ensure(cap);
      B200_CUDA(cudaMemsetAsync(h.d_rowdone.p, 0, (size_t)n * sizeof(int), st));
      B200_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int), st));
      B200_CUDA(cudaMemsetAsync(top.p, 0, sizeof(unsigned long long), st));
      B200_CUDA(cudaMemsetAsync(&h.ctrl.p->spin_timeout, 0, sizeof(int), st));
      int N = n; double tol = h.ilut_tol; const int *rows = h.d_rows.p, *cols = h.d_cols.p;
      int *pc = pcols.p, *rl = rlen.p, *rd = rdiag.p, *done = h.d_rowdone.p, *of = ovf.p; double *pv = pvals.p; long long *rsx = rstart.p;
      unsigned long long *tp = top.p; Ctrl *ctrl = h.ctrl.p;
      void *argv[] = {(void *)&N, (void *)&rows, (void *)&cols, (void *)&src, (void *)&tol, (void *)&pc, (void *)&pv, (void *)&cap, (void *)&tp,
                      (void *)&rsx, (void *)&rl, (void *)&rd, (void *)&done, (void *)&ctrl, (void *)&of};
      B200_CUDA(cudaLaunchCooperativeKernel((const void *)k_ilut_factor, dim3(blocks), dim3(ILUT_WARPS * 32), argv, smem, st));
      B200_CUDA(cudaMemcpyAsync(&ov, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaMemcpyAsync(h.h_ctrl, h.ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaStreamSynchronize(st));
      B200_REQUIRE(h.h_ctrl->spin_timeout == 0, "ILUT factorisation: dependency wait timed out");
      B200_REQUIRE(ov != 2, "ILUT: a row of the factor exceeds 1024 entries (tolerance too small for the accelerated path)");
      if (ov == 0) break;
      cap *= 4;                                                     // the pool was too small: again with four times the room
    }
    B200_REQUIRE(ov == 0, "ILUT: the factor does not fit in the entry pool");
    // the factor's own pattern: row pointers on the host (the plans are built there), columns / values / diagonal positions gathered on the device
    std::vector<int> hlen(n), hdg(n);
    B200_CUDA(cudaMemcpy(hlen.data(), rlen.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemcpy(hdg.data(), rdiag.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    h.hl_rows.assign((size_t)n + 1, 0); h.hl_diag.assign(n, 0);
    long long tot = 0;
    for (int i = 0; i < n; ++i) {
      B200_REQUIRE(hdg[i] >= 0, "ILUT: a row of the matrix has no diagonal entry");
      h.hl_rows[i] = (int)tot; h.hl_diag[i] = (int)tot + hdg[i]; tot += hlen[i];
      B200_REQUIRE(tot < 2147483647LL, "ILUT: the factor exceeds int32 entries");
    }
    h.hl_rows[n] = (int)tot;
    h.ilu_nnz = tot;
    h.dl_rows.ensure((size_t)n + 1); h.dl_cols.ensure((size_t)tot); h.dl_diag.ensure(n); h.d_ilu.ensure((size_t)tot);
    B200_CUDA(cudaMemcpyAsync(h.dl_rows.p, h.hl_rows.data(), ((size_t)n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    k_ilut_gather<<<NUM_SMS * 8, 256, 0, st>>>(n, h.dl_rows.p, rstart.p, rdiag.p, pcols.p, pvals.p, h.dl_cols.p, h.d_ilu.p, h.dl_diag.p);
    B200_CUDA(cudaGetLastError());
    h.hl_cols.resize((size_t)tot);
    B200_CUDA(cudaMemcpyAsync(h.hl_cols.data(), h.dl_cols.p, (size_t)tot * sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    pcols.release(); pvals.release(); rlen.release(); rdiag.release(); rstart.release(); ovf.release(); top.release();
    h.ilu_pat_ready = true;
    tri_analyse(h);                                                 // levels and sweep plans of THIS factor's pattern
    int eb = std::min((n + 255) / 256, NUM_SMS * 8);
    k_ilu0_invert_diag<<<eb, 256, 0, st>>>(n, h.d_ldiag(), h.d_ilu.p);            // 4323-4329
    sell_refresh_values(h, h.L, h.d_ilu.p);
    sell_refresh_values(h, h.U, h.d_ilu.p);
    if (h.U.nslots) k_gather_diag_slots<<<(h.U.nslots + 255) / 256, 256, 0, st>>>(h.U.nslots, h.U.perm.p, h.d_ldiag(), h.d_ilu.p, h.d_dinv_slot.p);
    B200_CUDA(cudaGetLastError());
  }
  B200_CUDA(cudaEventRecord(h.evf1, st));
  B200_CUDA(cudaStreamSynchronize(st));
  float ms = 0; B200_CUDA(cudaEventElapsedTime(&ms, h.evf0, h.evf1));
  h.st_factor_ms = ms;
  h.ilu_valid = true; h.ilu_exists = true;
  h.st_factor_launch = n > 0 ? 6 : 0;
  h.tri_mode = 0;
}

void ilu0_factor(Handle &h) {
  B200_REQUIRE(h.have_vals, "ILU0 requested before b200_set_values");
  if (h.ilut) { ilut_factor(h); return; }
  if (h.cholesky) { ichol_factor(h); return; }
  tri_analyse(h);

  if (h.tri_mode != 0 && h.tri_mode != 3 && h.tri_mode != 4) h.tri_mode = -2;     // anything else: the default
  if (h.tri_mode == 3) wave_analyse(h);
  else if (h.tri_mode == 4) lane_analyse(h);
  else if (h.tri_mode == -2) { wave_analyse(h); lane_analyse(h); }
  cudaStream_t st = h.stream;
  h.d_ilu.ensure(h.lnnz());
  B200_CUDA(cudaEventRecord(h.evf0, st));
  if (h.n > 0) {
    B200_CUDA(cudaMemsetAsync(h.d_rowdone.p, 0, (size_t)h.n * sizeof(int), st));
    B200_CUDA(cudaMemsetAsync(&h.ctrl.p->spin_timeout, 0, sizeof(int), st));
    const double *src = h.have_prec ? h.d_prec.p : h.d_vals.p;        // CRSMatrix.F90:3480-3484
    // position map (symbolic, once per structure / ILU order) when it fits: rows staged in shared memory, pivot rows of <= 64 upper
    // entries, at most 4 GB of one-byte positions; otherwise the kernel that searches on every update
    static const bool map_ok = !(getenv("B200_ILU_MAP") && atoi(getenv("B200_ILU_MAP")) == 0);
    if (map_ok && !h.ilu_map_tried) {
      h.ilu_map_tried = true;
      const std::vector<int> &R = h.lrows(), &Dg = h.ldiag();
      int maxu = 0, maxrow = 0; long long nlow_tot = 0;
      for (int i = 0; i < h.n; ++i) { maxu = std::max(maxu, R[i + 1] - Dg[i] - 1); maxrow = std::max(maxrow, R[i + 1] - R[i]); nlow_tot += Dg[i] - R[i]; }
      if (maxu >= 1 && maxu <= 64 && maxrow <= ILU_MAXROW && maxrow < 255 && nlow_tot * maxu <= (4LL << 30)) {
        std::vector<long long> pp((size_t)h.n + 1, 0);
        for (int i = 0; i < h.n; ++i) pp[i + 1] = pp[i] + (long long)(Dg[i] - R[i]) * maxu;
        h.d_ilu_posptr.ensure((size_t)h.n + 1); h.d_ilu_pos.ensure((size_t)std::max(1LL, pp[h.n]));
        B200_CUDA(cudaMemcpyAsync(h.d_ilu_posptr.p, pp.data(), ((size_t)h.n + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
        k_ilu_posmap<<<NUM_SMS * 8, 256, 0, st>>>(h.n, h.d_lrows(), h.d_lcols(), h.d_ldiag(), h.d_ilu_posptr.p, maxu, h.d_ilu_pos.p);
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaStreamSynchronize(st));                      // pp goes out of scope
        h.ilu_map_maxu = maxu;
      }
    }
    const bool use_map = map_ok && h.ilu_map_maxu > 0;
    // narrow rows (scalar 27-point stencils): the register kernel
    static const bool reg_ok = !(getenv("B200_ILU_REG") && atoi(getenv("B200_ILU_REG")) == 0);
    bool use_reg = use_map && reg_ok && h.ilu_map_maxu <= 32;
    if (use_reg) {
      if (h.ilu_reg_ok < 0) {
        const std::vector<int> &R = h.lrows(), &Dg = h.ldiag();
        bool ok = true;
        for (int i = 0; i < h.n && ok; ++i) ok = (R[i + 1] - R[i] <= 32) && (Dg[i] - R[i] <= 16);
        h.ilu_reg_ok = ok ? 1 : 0;
      }
      use_reg = h.ilu_reg_ok == 1;
    }
    const void *kern = use_reg ? (const void *)k_ilu0_factor_reg : use_map ? (const void *)k_ilu0_factor_map : (const void *)k_ilu0_factor;
    if (h.grid_ilu_kern != kern) { h.grid_ilu = persistent_blocks(kern, 256, 0); h.grid_ilu_kern = kern; }
    int blocks = std::max(1, std::min(h.grid_ilu, (h.L.nslots + 7) / 8));
    B200_CUDA(cudaEventRecord(h.evf0, st));                          // (the one-time symbolic map is not part of the factorisation time)
    if (use_map)
      launch_coresident(kern, blocks, 256, st, h.L.nslots, (const int *)h.L.perm.p, h.d_lrows(), h.d_lcols(), h.d_ldiag(), src,
                        (const int *)(h.ilu_sep() ? h.dl_src.p : nullptr), h.d_ilu.p, h.d_rowdone.p, h.ctrl.p, (const long long *)h.d_ilu_posptr.p, h.ilu_map_maxu,
                        (const unsigned char *)h.d_ilu_pos.p);
    else
      launch_coresident(kern, blocks, 256, st, h.L.nslots, (const int *)h.L.perm.p, h.d_lrows(), h.d_lcols(), h.d_ldiag(), src,
                        (const int *)(h.ilu_sep() ? h.dl_src.p : nullptr), h.d_ilu.p, h.d_rowdone.p, h.ctrl.p);
    int eb = std::min((h.n + 255) / 256, NUM_SMS * 8);
    k_ilu0_invert_diag<<<eb, 256, 0, st>>>(h.n, h.d_ldiag(), h.d_ilu.p);
    sell_refresh_values(h, h.L, h.d_ilu.p);
    sell_refresh_values(h, h.U, h.d_ilu.p);
    if (h.U.nslots) k_gather_diag_slots<<<(h.U.nslots + 255) / 256, 256, 0, st>>>(h.U.nslots, h.U.perm.p, h.d_ldiag(), h.d_ilu.p, h.d_dinv_slot.p);
    if (h.wv.ready) wave_refresh_values(h);
    if (h.lt.ready) lane_refresh_values(h);
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
  if (h.tri_mode == -2) tri_autotune_wave(h);
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
// Vectors of the solves live in slot (level) order: `out` has one entry per slot, the column ids of
// T are slot ids.  rhs_idx == nullptr: rhs is in natural order (forward sweep input); otherwise
// rhs[rhs_idx[slot]] (backward sweep reading the forward result).  nat_out != nullptr: the result is
// also scattered to natural order (the preconditioned vector handed back to the Krylov method).
// guarded load of a result-vector entry: predicated in PTX, no branch
__device__ __forceinline__ double ld_relaxed_pred(const double *p, bool pred, double old) {
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q ld.relaxed.gpu.global.f64 %0, [%1]; }" : "+d"(old) : "l"(p), "r"((unsigned)pred));
  return old;
}
// Rows of at most CH entries (one register chunk) take a path tuned for the critical dependency: what a
// row still has to do AFTER its last operand became visible decides the time per level.  The operands
// of the previous level are the entries next to the diagonal (natural FE numberings: the last TAIL
// entries of a lower row, the first TAIL of an upper row; L is stored right-aligned so that these are
// the same registers in every lane).  The other ("head") operands are normally there at the first
// read: their products are formed, and for L already subtracted, while the tail is still being
// polled; after the last arrival only TAIL multiply-subtract pairs (L) or TAIL pairs + the head
// subtractions (U: the sum must run left to right and its first term arrives last) remain.  Every
// guarded load is predicated and the loops are warp-uniform: no divergence bookkeeping on the path.
// The arithmetic is unchanged: p = v*x rounded, s = s - p rounded, left to right.
constexpr int TRI_TAIL = 6;
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
    auto gate = [&]() {                                   // throttle: wait for the wavefront to come near
      const int wl = slice_level[slice] - lookahead;
      if (lane == 0 && wl >= 0) {
        const int need = lvl_slices[wl];
        while (ld_relaxed_i(lvl_done + wl * 32) < need) {
          if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
          if (gate_sleep) __nanosleep(gate_sleep);
        }
      }
      __syncwarp();
    };
    if (W <= CH) {
      // register k <-> stored position k - sh; entry present iff lo <= k < hi
      const int sh = UPPER ? 0 : CH - W;
      const int lo = UPPER ? 0 : CH - len, hi = UPPER ? len : CH;
      int c[CH]; double v[CH], x[CH];
#pragma unroll
      for (int k = 0; k < CH; ++k) {                      // matrix entries: streamed from HBM ahead of the wavefront
        const int pos = min(max(k - sh, 0), max(W - 1, 0));
        c[k] = W ? ld_stream(cp + pos * 32) : 0;
        v[k] = W ? ld_stream(vp + pos * 32) : 0.0;
      }
      gate();
      unsigned has = 0;
#pragma unroll
      for (int k = 0; k < CH; ++k) has |= (unsigned)(k >= lo && k < hi) << k;
#pragma unroll
      for (int k = 0; k < CH; ++k) x[k] = ld_relaxed_pred(out + c[k], (has >> k) & 1u, 0.0);   // all gathers in flight
      constexpr unsigned TAILMASK = UPPER ? ((1u << TRI_TAIL) - 1u) : (((1u << TRI_TAIL) - 1u) << (CH - TRI_TAIL));
      // head operands
      unsigned pend = 0;
#pragma unroll
      for (int k = 0; k < CH; ++k) pend |= (unsigned)(is_sentinel(x[k])) << k;
      pend &= has;
      while (__any_sync(0xffffffffu, (pend & ~TAILMASK) != 0)) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        if (spin_sleep) __nanosleep(spin_sleep);
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          if (TAILMASK & (1u << k)) continue;
          x[k] = ld_relaxed_pred(out + c[k], (pend >> k) & 1u, x[k]);
          if (!is_sentinel(x[k])) pend &= ~(1u << k);
        }
      }
      double p[CH];
#pragma unroll
      for (int k = 0; k < CH; ++k) p[k] = __dmul_rn(v[k], x[k]);
      if (!UPPER) {
#pragma unroll
        for (int k = 0; k < CH - TRI_TAIL; ++k) { const double t = __dsub_rn(s, p[k]); s = (has >> k) & 1u ? t : s; }
      }
      // tail operands: the previous level
#pragma unroll
      for (int k = 0; k < CH; ++k) if (TAILMASK & (1u << k)) { x[k] = ld_relaxed_pred(out + c[k], (pend >> k) & 1u, x[k]); if (!is_sentinel(x[k])) pend &= ~(1u << k); }
      while (__any_sync(0xffffffffu, pend != 0)) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        if (spin_sleep) __nanosleep(spin_sleep);
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          if (!(TAILMASK & (1u << k))) continue;
          x[k] = ld_relaxed_pred(out + c[k], (pend >> k) & 1u, x[k]);
          if (!is_sentinel(x[k])) pend &= ~(1u << k);
        }
      }
      if (!UPPER) {
#pragma unroll
        for (int k = CH - TRI_TAIL; k < CH; ++k) { const double t = __dsub_rn(s, __dmul_rn(v[k], x[k])); s = (has >> k) & 1u ? t : s; }
      } else {
#pragma unroll
        for (int k = 0; k < TRI_TAIL; ++k) { const double t = __dsub_rn(s, __dmul_rn(v[k], x[k])); s = (has >> k) & 1u ? t : s; }
#pragma unroll
        for (int k = TRI_TAIL; k < CH; ++k) { const double t = __dsub_rn(s, p[k]); s = (has >> k) & 1u ? t : s; }
      }
    } else {
      const int first = UPPER ? 0 : W - len, last = UPPER ? len : W;
      bool gated = false;
      for (int j0 = 0; j0 < W; j0 += CH) {
        int c[CH]; double v[CH], xv[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          const int pos = min(j0 + k, W - 1);
          c[k] = ld_stream(cp + pos * 32);
          v[k] = ld_stream(vp + pos * 32);
        }
        if (!gated) { gate(); gated = true; }
        unsigned has = 0;
#pragma unroll
        for (int k = 0; k < CH; ++k) has |= (unsigned)(j0 + k >= first && j0 + k < last) << k;
#pragma unroll
        for (int k = 0; k < CH; ++k) xv[k] = ld_relaxed_pred(out + c[k], (has >> k) & 1u, 0.0);
        unsigned pend = 0;
#pragma unroll
        for (int k = 0; k < CH; ++k) pend |= (unsigned)(is_sentinel(xv[k])) << k;
        pend &= has;
        while (__any_sync(0xffffffffu, pend != 0)) {
          if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
          if (spin_sleep) __nanosleep(spin_sleep);
#pragma unroll
          for (int k = 0; k < CH; ++k) { xv[k] = ld_relaxed_pred(out + c[k], (pend >> k) & 1u, xv[k]); if (!is_sentinel(xv[k])) pend &= ~(1u << k); }
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) { const double t = nfms(s, v[k], xv[k]); s = (has >> k) & 1u ? t : s; }
      }
      if (!gated) gate();
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

// Rows of 17 .. 16*NCH entries (3-dof elasticity: ~40 per triangle; ILU(1) of a 27-point stencil: ~31).  The generic chunk loop
// of k_sptrsv pays an HBM + L2 round trip per chunk AFTER the chunk before it is complete, and the upper sweep starts with
// the operands that arrive last: 2.6 us per level on the elasticity operand.  Here the TAIL (the TRI_TAIL entries next to
// the diagonal = the previous level) is requested first and polled last; the head chunks are consumed while the tail is
// still in flight -- subtracted straight away for L (they precede the tail in column order), kept as rounded products
// for U (they follow it, and the sum must run left to right) -- so that after the last arrival only the tail pairs and,
// for U, a chain of subtractions remain.  One block per SM: the product registers need the whole register file.
template <bool UPPER, int NCH>
__global__ void __launch_bounds__(256, 1) k_sptrsv_wide(SellView T, const int *__restrict__ slice_level, const int *__restrict__ lvl_slices,
                                                         int *lvl_done, int lookahead, unsigned gate_sleep, unsigned spin_sleep,
                                                         const double *__restrict__ dinv_slot, const double *__restrict__ rhs,
                                                         const int *__restrict__ rhs_idx, double *out, double *__restrict__ nat_out,
                                                         Ctrl *ctrl) {
  if (ctrl->done) return;
  constexpr int WT = 16 * NCH;
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
    // register position k <-> stored position k - sh; entry present iff lo <= k < hi
    const int sh = UPPER ? 0 : WT - W;
    const int lo = UPPER ? 0 : WT - len, hi = UPPER ? len : WT;
    const int wl = max(W - 1, 0);
    constexpr int K0 = UPPER ? 0 : WT - TRI_TAIL;                    // first tail position
    int ct[TRI_TAIL]; double vt[TRI_TAIL], xt[TRI_TAIL];
#pragma unroll
    for (int t = 0; t < TRI_TAIL; ++t) {
      const int pos = min(max(K0 + t - sh, 0), wl);
      ct[t] = W ? ld_stream(cp + pos * 32) : 0;
      vt[t] = W ? ld_stream(vp + pos * 32) : 0.0;
    }
    {                                                             // throttle: wait for the wavefront to come near
      const int wlv = slice_level[slice] - lookahead;
      if (lane == 0 && wlv >= 0) {
        const int need = lvl_slices[wlv];
        while (ld_relaxed_i(lvl_done + wlv * 32) < need) {
          if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
          if (gate_sleep) __nanosleep(gate_sleep);
        }
      }
      __syncwarp();
    }
    unsigned tpend = 0;
#pragma unroll
    for (int t = 0; t < TRI_TAIL; ++t) {
      const bool has = (K0 + t >= lo) && (K0 + t < hi);
      xt[t] = ld_relaxed_pred(out + ct[t], has, 0.0);
      tpend |= (unsigned)has << t;
    }
    const unsigned thas = tpend;
    double p[UPPER ? WT : 1];
#pragma unroll
    for (int k = 0; k < (UPPER ? WT : 1); ++k) p[k] = 0.0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      if (UPPER ? (ch * 16 >= W) : (ch * 16 + 15 < WT - W)) continue;   // (warp-uniform) nothing stored in this chunk
      int c[16]; double v[16], x[16];
      unsigned has = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = ch * 16 + j;
        const bool tail = UPPER ? (k < TRI_TAIL) : (k >= WT - TRI_TAIL);
        const int pos = min(max(k - sh, 0), wl);
        c[j] = W ? ld_stream(cp + pos * 32) : 0;
        v[j] = W ? ld_stream(vp + pos * 32) : 0.0;
        has |= (unsigned)(!tail && k >= lo && k < hi) << j;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = ld_relaxed_pred(out + c[j], (has >> j) & 1u, 0.0);
      unsigned pend = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) pend |= (unsigned)(is_sentinel(x[j])) << j;
      pend &= has;
      while (__any_sync(0xffffffffu, pend != 0)) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        if (spin_sleep) __nanosleep(spin_sleep);
#pragma unroll
        for (int j = 0; j < 16; ++j) { x[j] = ld_relaxed_pred(out + c[j], (pend >> j) & 1u, x[j]); if (!is_sentinel(x[j])) pend &= ~(1u << j); }
      }
      if (!UPPER) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { const double t = __dsub_rn(s, __dmul_rn(v[j], x[j])); s = (has >> j) & 1u ? t : s; }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) p[ch * 16 + j] = (has >> j) & 1u ? __dmul_rn(v[j], x[j]) : 0.0;
      }
    }
    // tail operands: the previous level
#pragma unroll
    for (int t = 0; t < TRI_TAIL; ++t) if (!is_sentinel(xt[t])) tpend &= ~(1u << t);
    while (__any_sync(0xffffffffu, tpend != 0)) {
      if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
      if (spin_sleep) __nanosleep(spin_sleep);
#pragma unroll
      for (int t = 0; t < TRI_TAIL; ++t) { xt[t] = ld_relaxed_pred(out + ct[t], (tpend >> t) & 1u, xt[t]); if (!is_sentinel(xt[t])) tpend &= ~(1u << t); }
    }
#pragma unroll
    for (int t = 0; t < TRI_TAIL; ++t) { const double u = __dsub_rn(s, __dmul_rn(vt[t], xt[t])); s = (thas >> t) & 1u ? u : s; }
    if (UPPER) {
#pragma unroll
      for (int k = TRI_TAIL; k < WT; ++k) { const double u = __dsub_rn(s, p[k]); s = (k < hi) ? u : s; }   // absent entries hold +0.0 and are skipped
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
// The same sweep on the node-lane layout (structure.cu node_lane_layout: ND interleaved dofs per node, lane = ni * ND + d).  The rows
// of a node are solved in ONE pass of a warp: everything a row needs from OTHER nodes is gathered / polled exactly as in k_sptrsv_wide
// (all lanes in parallel), what it needs from its own node -- the entries adjacent to the diagonal: the last d entries of a lower row,
// the first ND-1-d of an upper row -- is passed between lanes by shuffle in ND short rounds.  One L2 hand-off per NODE level instead of
// one per row level.  Arithmetic: the reference's left-to-right order (own-node columns come last in a lower row and first in an upper
// row, which is why the upper rows of a node are summed one after the other); the value another row receives by shuffle is the
// canonicalised one it would have read from memory.  Bit-identical to k_sptrsv / CRS_LUSolve.
template <bool UPPER, int NCH, int ND>
__global__ void __launch_bounds__(256, 1) k_sptrsv_wide_node(SellView T, const int *__restrict__ slice_level, const int *__restrict__ lvl_slices,
                                                              int *lvl_done, int lookahead, unsigned gate_sleep, unsigned spin_sleep,
                                                              const double *__restrict__ dinv_slot, const double *__restrict__ rhs,
                                                              const int *__restrict__ rhs_idx, double *out, double *__restrict__ nat_out,
                                                              Ctrl *ctrl) {
  if (ctrl->done) return;
  constexpr int WT = 16 * NCH;
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int ni = lane / ND, d = lane - ni * ND, nbase = ni * ND;     // node of the lane inside the slice, dof, first lane of the node
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
    const double r0 = row >= 0 ? rhs[rhs_idx ? rhs_idx[slot] : row] : 0.0;
    const double dinv = (UPPER && row >= 0) ? dinv_slot[slot] : 1.0;
    long long spins = 0;
    const int sh = UPPER ? 0 : WT - W;
    const int lo = UPPER ? 0 : WT - len, hi = UPPER ? len : WT;
    const int wl = max(W - 1, 0);
    constexpr int K0 = UPPER ? 0 : WT - TRI_TAIL;                    // first tail position
    int ct[TRI_TAIL]; double vt[TRI_TAIL], xt[TRI_TAIL];
#pragma unroll
    for (int t = 0; t < TRI_TAIL; ++t) {
      const int pos = min(max(K0 + t - sh, 0), wl);
      ct[t] = W ? ld_stream(cp + pos * 32) : 0;
      vt[t] = W ? ld_stream(vp + pos * 32) : 0.0;
    }
    {                                                             // throttle: wait for the wavefront to come near
      const int wlv = slice_level[slice] - lookahead;
      if (lane == 0 && wlv >= 0) {
        const int need = lvl_slices[wlv];
        while (ld_relaxed_i(lvl_done + wlv * 32) < need) {
          if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
          if (gate_sleep) __nanosleep(gate_sleep);
        }
      }
      __syncwarp();
    }
    // tail entries that belong to the node itself: lower row, the last d ; upper row, the first ND-1-d
    unsigned thas = 0, town = 0;
#pragma unroll
    for (int t = 0; t < TRI_TAIL; ++t) {
      const bool has = (K0 + t >= lo) && (K0 + t < hi);
      const bool own = has && row >= 0 && (UPPER ? (t < ND - 1 - d) : (t >= TRI_TAIL - d));
      thas |= (unsigned)has << t; town |= (unsigned)own << t;
      xt[t] = ld_relaxed_pred(out + ct[t], has && !own, 0.0);
    }
    unsigned tpend = thas & ~town;
    double s = r0;
    double p[UPPER ? WT : 1];
#pragma unroll
    for (int k = 0; k < (UPPER ? WT : 1); ++k) p[k] = 0.0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      if (UPPER ? (ch * 16 >= W) : (ch * 16 + 15 < WT - W)) continue;   // (warp-uniform) nothing stored in this chunk
      int c[16]; double v[16], x[16];
      unsigned has = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = ch * 16 + j;
        const bool tail = UPPER ? (k < TRI_TAIL) : (k >= WT - TRI_TAIL);
        const int pos = min(max(k - sh, 0), wl);
        c[j] = W ? ld_stream(cp + pos * 32) : 0;
        v[j] = W ? ld_stream(vp + pos * 32) : 0.0;
        has |= (unsigned)(!tail && k >= lo && k < hi) << j;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = ld_relaxed_pred(out + c[j], (has >> j) & 1u, 0.0);
      unsigned pend = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) pend |= (unsigned)(is_sentinel(x[j])) << j;
      pend &= has;
      while (__any_sync(FULL, pend != 0)) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        if (spin_sleep) __nanosleep(spin_sleep);
#pragma unroll
        for (int j = 0; j < 16; ++j) { x[j] = ld_relaxed_pred(out + c[j], (pend >> j) & 1u, x[j]); if (!is_sentinel(x[j])) pend &= ~(1u << j); }
      }
      if (!UPPER) {
#pragma unroll
        for (int j = 0; j < 16; ++j) { const double t = __dsub_rn(s, __dmul_rn(v[j], x[j])); s = (has >> j) & 1u ? t : s; }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) p[ch * 16 + j] = (has >> j) & 1u ? __dmul_rn(v[j], x[j]) : 0.0;
      }
    }
    // tail operands of OTHER nodes: the previous level
#pragma unroll
    for (int t = 0; t < TRI_TAIL; ++t) if (!is_sentinel(xt[t])) tpend &= ~(1u << t);
    while (__any_sync(FULL, tpend != 0)) {
      if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
      if (spin_sleep) __nanosleep(spin_sleep);
#pragma unroll
      for (int t = 0; t < TRI_TAIL; ++t) { xt[t] = ld_relaxed_pred(out + ct[t], (tpend >> t) & 1u, xt[t]); if (!is_sentinel(xt[t])) tpend &= ~(1u << t); }
    }
    double res = 0.0;
    if (!UPPER) {
      // the tail entries of other nodes, then the node's own dofs 0 .. d-1 (column order), one shuffle round per dof
#pragma unroll
      for (int t = 0; t < TRI_TAIL; ++t) { const double u = __dsub_rn(s, __dmul_rn(vt[t], xt[t])); s = ((thas & ~town) >> t) & 1u ? u : s; }
#pragma unroll
      for (int r = 0; r < ND - 1; ++r) {
        double fin = s;                                           // final for the lanes with d <= r
        if (fin != fin) fin = __longlong_as_double((long long)CANON_NAN);
        const double xr = __shfl_sync(FULL, fin, nbase + r);
        // own entry of dof r sits at tail index TRI_TAIL - d + r
        double vo = 0.0; bool ho = false;
#pragma unroll
        for (int t = 0; t < TRI_TAIL; ++t) if (t == TRI_TAIL - d + r) { vo = vt[t]; ho = (town >> t) & 1u; }
        if (d > r && ho) s = __dsub_rn(s, __dmul_rn(vo, xr));
      }
      res = s;
    } else {
      // upper rows of a node one after the other, dof ND-1 first: own dofs d+1 .. ND-1 (column order), other nodes' tail, the products
      double xo[ND];
#pragma unroll
      for (int r = 0; r < ND; ++r) xo[r] = 0.0;
#pragma unroll
      for (int r = ND - 1; r >= 0; --r) {
        double a = r0;
#pragma unroll
        for (int t = 0; t < TRI_TAIL; ++t) {
          double xv = xt[t];
          if ((town >> t) & 1u) {                                 // own entry: dof d + 1 + t
            xv = xo[ND - 1];
#pragma unroll
            for (int q = 0; q < ND - 1; ++q) if (d + 1 + t == q) xv = xo[q];
          }
          const double u = __dsub_rn(a, __dmul_rn(vt[t], xv));
          a = (thas >> t) & 1u ? u : a;
        }
#pragma unroll
        for (int k = TRI_TAIL; k < WT; ++k) { const double u = __dsub_rn(a, p[k]); a = (k < hi) ? u : a; }
        double fin = __dmul_rn(dinv, a);
        if (fin != fin) fin = __longlong_as_double((long long)CANON_NAN);
        if (d == r) res = fin;
        xo[r] = __shfl_sync(FULL, fin, nbase + r);
      }
    }
    if (row >= 0) {
      if (res != res) res = __longlong_as_double((long long)CANON_NAN);
      st_relaxed(out + slot, res);
      if (nat_out) nat_out[row] = res;
    }
    __syncwarp();
    if (lane == 0) atomicAdd(lvl_done + slice_level[slice] * 32, 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Symmetric Gauss-Seidel sweeps of itermethod_sgs (IterativeMethods.F90:219-283): forward i = 1..n, then backward i = n..1,
//   s = sum_j A_ij x_j (current x, left to right) ;  x_i = x_i + Omega * (b_i - s) / a_ii.
// The same wavefront as the triangular solves, on the matrix itself: rows in dependency-level order (the L / U plans of the
// matrix pattern), a row's already-updated operands (j < i forward, j > i backward) are polled from the sentinel-filled
// output vector, the others come from the input vector.  One lane walks one CRS row, so the sum runs in column order with
// separate roundings exactly as the reference loop.
template <bool BACKWARD>
__global__ void __launch_bounds__(256, 2) k_sgs_sweep(int nslices, const int *__restrict__ perm, const int *__restrict__ slice_level,
                                                       const int *__restrict__ lvl_slices, int *lvl_done, int lookahead, unsigned gate_sleep,
                                                       const int *__restrict__ rows, const int *__restrict__ cols, const int *__restrict__ diag,
                                                       const double *__restrict__ vals, const double *__restrict__ b, double omega,
                                                       const double *__restrict__ xin, double *xout, Ctrl *ctrl) {
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int slice = gwarp; slice < nslices; slice += nwarps) {
    const int r = perm[slice * 32 + lane];
    long long spins = 0;
    const int wl = slice_level[slice] - lookahead;
    if (lane == 0 && wl >= 0) {
      const int need = lvl_slices[wl];
      while (ld_relaxed_i(lvl_done + wl * 32) < need) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        if (gate_sleep) __nanosleep(gate_sleep);
      }
    }
    __syncwarp();
    if (r >= 0) {
      double s = 0.0;
      for (int p = rows[r]; p < rows[r + 1]; ++p) {
        const int c = cols[p];
        double xv;
        if (BACKWARD ? (c > r) : (c < r)) {
          xv = ld_relaxed(xout + c);
          while (is_sentinel(xv)) {
            if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
            xv = ld_relaxed(xout + c);
          }
        } else xv = xin[c];
        s = nfma(s, xv, vals[p]);
      }
      double xn = __dadd_rn(xin[r], __ddiv_rn(__dmul_rn(omega, __dsub_rn(b[r], s)), vals[diag[r]]));
      if (xn != xn) xn = __longlong_as_double((long long)CANON_NAN);
      st_relaxed(xout + r, xn);
    }
    __syncwarp();
    if (lane == 0) atomicAdd(lvl_done + slice_level[slice] * 32, 1);
  }
}
__global__ void k_sgs_prepare(int n, double *a, double *b2, int *counters, int ncounters) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, st = gridDim.x * blockDim.x;
  for (int i = i0; i < n; i += st) { a[i] = sent; b2[i] = sent; }
  for (int i = i0; i < ncounters; i += st) counters[i * 32] = 0;
}
// x <- one forward and one backward sweep; t1, t2: work vectors of length n
void sgs_sweeps(Handle &h, const double *b, double *x, double *t1, double *t2, double omega) {
  B200_REQUIRE(h.have_vals, "SGS before b200_set_values");
  if (h.n == 0) return;
  if (h.ilu_sep()) { h.ilu_order = 0; h.bilu_blocks = 0; ilu_invalidate(h); }     // the sweeps need the plans of the matrix pattern itself
  if (h.tri_node) { h.tri_node_off = true; ilu_invalidate(h); }                   // ... in the row-level layout (k_sgs_sweep polls per row)
  tri_analyse(h);
  cudaStream_t st = h.stream;
  static int grid_f = 0, grid_b = 0;
  if (!grid_f) {
    grid_f = persistent_blocks((const void *)k_sgs_sweep<false>, 256, 0);
    grid_b = persistent_blocks((const void *)k_sgs_sweep<true>, 256, 0);
  }
  k_sgs_prepare<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, st>>>(h.n, t1, t2, h.tri_counters.p, h.nlev_f + h.nlev_b + 2);
  const int la = h.tri_lookahead; const unsigned gs = h.tri_gate_sleep;
  const int bf = std::max(1, std::min(grid_f, (h.L.nslices + 7) / 8)), bb = std::max(1, std::min(grid_b, (h.U.nslices + 7) / 8));
  launch_coresident((const void *)k_sgs_sweep<false>, bf, 256, st, h.L.nslices, (const int *)h.L.perm.p, (const int *)h.L.gate.p, (const int *)h.d_lvlcnt_f.p,
                    h.tri_counters.p, la, gs, (const int *)h.d_rows.p, (const int *)h.d_cols.p, (const int *)h.d_diag.p, (const double *)h.d_vals.p, b, omega,
                    (const double *)x, t1, h.ctrl.p);
  launch_coresident((const void *)k_sgs_sweep<true>, bb, 256, st, h.U.nslices, (const int *)h.U.perm.p, (const int *)h.U.gate.p, (const int *)h.d_lvlcnt_b.p,
                    h.tri_counters.p + (size_t)(h.nlev_f + 1) * 32, la, gs, (const int *)h.d_rows.p, (const int *)h.d_cols.p, (const int *)h.d_diag.p,
                    (const double *)h.d_vals.p, b, omega, (const double *)t1, t2, h.ctrl.p);
  B200_CUDA(cudaMemcpyAsync(x, t2, (size_t)h.n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  B200_CUDA(cudaGetLastError());
  h.st_launch += 3;
}

__global__ void k_tri_prepare(int na, double *a, int nb, double *b, int *counters, int ncounters) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, st = gridDim.x * blockDim.x;
  for (int i = i0; i < na; i += st) a[i] = sent;
  for (int i = i0; i < nb; i += st) b[i] = sent;
  for (int i = i0; i < ncounters; i += st) counters[i * 32] = 0;
}

// Incomplete Cholesky solve, CRS_LUSolve with A % Cholesky (CRSMatrix.F90:4618-4638), as two wavefront sweeps over natural-order vectors
// (k_sgs_sweep's scheme: rows in dependency-level order, operands polled from the sentinel-filled result vector, one lane per row so that
// every sum runs in the reference's order with separate roundings).
//   forward  (4621-4628): z_i = (b_i - sum_{j<i} L_ij z_j) * Linv_ii, columns ascending;
//   backward (4632-4638): x_c = (z_c - sum_{i>c} L_ic x_i) * Linv_cc, rows i DESCENDING (the order in which the column-oriented loop
//                         reaches b(c)); the lists come from ichol_analyse.
template <bool BACKWARD>
__global__ void __launch_bounds__(256, 2) k_ichol_sweep(int nslices, const int *__restrict__ perm, const int *__restrict__ slice_level,
                                                         const int *__restrict__ lvl_slices, int *lvl_done, int lookahead, unsigned gate_sleep,
                                                         const int *__restrict__ ptr, const int *__restrict__ idx, const int *__restrict__ pos,
                                                         const int *__restrict__ diag, const double *__restrict__ LU, const double *__restrict__ rhs,
                                                         double *out, double *copy, Ctrl *ctrl) {
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int slice = gwarp; slice < nslices; slice += nwarps) {
    const int r = perm[slice * 32 + lane];
    long long spins = 0;
    const int wl = slice_level[slice] - lookahead;
    if (lane == 0 && wl >= 0) {
      const int need = lvl_slices[wl];
      while (ld_relaxed_i(lvl_done + wl * 32) < need) {
        if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
        if (gate_sleep) __nanosleep(gate_sleep);
      }
    }
    __syncwarp();
    if (r >= 0) {
      double s = rhs[r];
      const int p0 = BACKWARD ? ptr[r] : ptr[r], p1 = BACKWARD ? ptr[r + 1] : diag[r];     // forward: ptr = Rows, the strict lower part
      for (int p = p0; p < p1; ++p) {
        const int c = idx[p];
        double xv = ld_relaxed(out + c);
        while (is_sentinel(xv)) {
          if (++spins > SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
          xv = ld_relaxed(out + c);
        }
        s = nfms(s, LU[BACKWARD ? pos[p] : p], xv);
      }
      double xn = __dmul_rn(s, LU[diag[r]]);
      if (xn != xn) xn = __longlong_as_double((long long)CANON_NAN);
      st_relaxed(out + r, xn);
      if (copy) copy[r] = xn;
    }
    __syncwarp();
    if (lane == 0) atomicAdd(lvl_done + slice_level[slice] * 32, 1);
  }
}
__global__ void k_ichol_prepare(int n, double *a, double *b2, int *counters, int ncounters) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x, st = gridDim.x * blockDim.x;
  for (int i = i0; i < n; i += st) { a[i] = sent; b2[i] = sent; }
  for (int i = i0; i < ncounters; i += st) counters[i * 32] = 0;
}
static void lu_apply_ichol(Handle &h, double *u, const double *v) {
  B200_REQUIRE(h.ch_ready, "incomplete Cholesky solve without a plan");
  cudaStream_t st = h.stream;
  static int grid_f = 0, grid_b = 0;
  if (!grid_f) {
    grid_f = persistent_blocks((const void *)k_ichol_sweep<false>, 256, 0);
    grid_b = persistent_blocks((const void *)k_ichol_sweep<true>, 256, 0);
  }
  k_ichol_prepare<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, st>>>(h.n, h.ch_y.p, h.ch_x.p, h.ch_counters.p, h.nlev_f + h.ch_nlev + 2);
  const int la = h.tri_lookahead; const unsigned gs = h.tri_gate_sleep;
  const int bf = std::max(1, std::min(grid_f, (h.L.nslices + 7) / 8)), bb = std::max(1, std::min(grid_b, (h.ch_nslices + 7) / 8));
  launch_coresident((const void *)k_ichol_sweep<false>, bf, 256, st, h.L.nslices, (const int *)h.L.perm.p, (const int *)h.L.gate.p, (const int *)h.d_lvlcnt_f.p,
                    h.ch_counters.p, la, gs, h.d_lrows(), h.d_lcols(), (const int *)nullptr, h.d_ldiag(), (const double *)h.d_ilu.p, v, h.ch_y.p,
                    (double *)nullptr, h.ctrl.p);
  launch_coresident((const void *)k_ichol_sweep<true>, bb, 256, st, h.ch_nslices, (const int *)h.ch_perm.p, (const int *)h.ch_gate.p, (const int *)h.ch_lvlcnt.p,
                    h.ch_counters.p + (size_t)(h.nlev_f + 1) * 32, la, gs, (const int *)h.ch_ptr.p, (const int *)h.ch_row.p, (const int *)h.ch_pos.p, h.d_ldiag(),
                    (const double *)h.d_ilu.p, (const double *)h.ch_y.p, h.ch_x.p, u, h.ctrl.p);
  B200_CUDA(cudaGetLastError());
  h.st_launch += 3; h.st_pcond++;
}

static void lu_launch_kernels(Handle &h, const void *kl, const void *ku, double *u, const double *v) {
  cudaStream_t st = h.stream;
  if (!h.grid_tri_l) {
    h.grid_tri_l = persistent_blocks(kl, 256, h.tri_blocks_per_sm);
    h.grid_tri_u = persistent_blocks(ku, 256, h.tri_blocks_per_sm);
  }
  int bl = std::max(1, std::min(h.grid_tri_l, (h.L.nslices + 7) / 8));
  int bu = std::max(1, std::min(h.grid_tri_u, (h.U.nslices + 7) / 8));
  const int la = h.tri_lookahead; const unsigned gs = h.tri_gate_sleep, ss = h.tri_spin_sleep;
  launch_coresident(kl, bl, 256, st, h.L.view(), (const int *)h.L.gate.p, (const int *)h.d_lvlcnt_f.p, h.tri_counters.p, la, gs, ss,
                    (const double *)nullptr, v, (const int *)nullptr, h.d_yl.p, (double *)nullptr, h.ctrl.p);
  launch_coresident(ku, bu, 256, st, h.U.view(), (const int *)h.U.gate.p, (const int *)h.d_lvlcnt_b.p,
                    h.tri_counters.p + (size_t)(h.nlev_f + 1) * 32, la, gs, ss,
                    (const double *)h.d_dinv_slot.p, (const double *)h.d_yl.p, (const int *)h.d_urhs.p, h.d_xu.p, u, h.ctrl.p);
}

void lu_apply(Handle &h, double *u, const double *v) {
  B200_REQUIRE(h.ilu_valid, "LU preconditioner applied without a valid ILU0 factor");
  if (h.n == 0) return;
  if (h.cholesky) { lu_apply_ichol(h, u, v); return; }
  if (h.tri_mode == 3 && h.wv.ready) { lu_apply_wave(h, u, v); return; }   // grid stencils; not detected -> level kernel
  if (h.tri_mode == 4 && h.lt.ready) { lu_apply_lane(h, u, v); return; }   // grid stencils; not detected -> level kernel
  k_tri_prepare<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, h.stream>>>(h.L.nslots, h.d_yl.p, h.U.nslots, h.d_xu.p, h.tri_counters.p, h.nlev_f + h.nlev_b + 2);
  if (h.tri_node >= 2) {                                           // node-lane plans: only the node-aware sweeps may run on them
    const int nch = (h.tri_maxw + 15) / 16;
    B200_REQUIRE(nch >= 1 && nch <= 4, "node-lane plan with rows wider than 64 entries");
    const void *kl = nullptr, *ku = nullptr;
#define B200_NODE_CASE(ND_)                                                                                                       \
    case ND_:                                                                                                                       \
      kl = nch == 1 ? (const void *)k_sptrsv_wide_node<false, 1, ND_> : nch == 2 ? (const void *)k_sptrsv_wide_node<false, 2, ND_>    \
         : nch == 3 ? (const void *)k_sptrsv_wide_node<false, 3, ND_> : (const void *)k_sptrsv_wide_node<false, 4, ND_>;             \
      ku = nch == 1 ? (const void *)k_sptrsv_wide_node<true, 1, ND_> : nch == 2 ? (const void *)k_sptrsv_wide_node<true, 2, ND_>      \
         : nch == 3 ? (const void *)k_sptrsv_wide_node<true, 3, ND_> : (const void *)k_sptrsv_wide_node<true, 4, ND_>;               \
      break;
    switch (h.tri_node) {
      B200_NODE_CASE(2) B200_NODE_CASE(3) B200_NODE_CASE(4) B200_NODE_CASE(5) B200_NODE_CASE(6)
      default: B200_REQUIRE(false, "node-lane plan: unsupported dofs per node");
    }
#undef B200_NODE_CASE
    if (!h.tri_node_u) {                                            // backward plan in the row-level layout: the row-per-thread sweeps
      if (h.tri_maxw <= 8) ku = (const void *)k_sptrsv<true, 8>;
      else if (h.tri_maxw <= 16 || h.tri_maxw > 48) ku = (const void *)k_sptrsv<true, 16>;
      else if (h.tri_maxw <= 32) ku = (const void *)k_sptrsv_wide<true, 2>;
      else ku = (const void *)k_sptrsv_wide<true, 3>;
    }
    lu_launch_kernels(h, kl, ku, u, v);
    B200_CUDA(cudaGetLastError());
    h.st_launch += 3; h.st_pcond++;
    return;
  }
  static const bool wide_ok = !(getenv("B200_TRI_WIDE") && atoi(getenv("B200_TRI_WIDE")) == 0);
  if (h.tri_maxw <= 8) lu_launch_kernels(h, (const void *)k_sptrsv<false, 8>, (const void *)k_sptrsv<true, 8>, u, v);
  else if (h.tri_maxw <= 16 || !wide_ok || h.tri_maxw > 48) lu_launch_kernels(h, (const void *)k_sptrsv<false, 16>, (const void *)k_sptrsv<true, 16>, u, v);
  else if (h.tri_maxw <= 32) lu_launch_kernels(h, (const void *)k_sptrsv_wide<false, 2>, (const void *)k_sptrsv_wide<true, 2>, u, v);
  else lu_launch_kernels(h, (const void *)k_sptrsv_wide<false, 3>, (const void *)k_sptrsv_wide<true, 3>, u, v);
  B200_CUDA(cudaGetLastError());
  h.st_launch += 3; h.st_pcond++;
}

// Default (B200_TRI_MODE unset): when the factor is that of a structured-grid stencil, the level kernel, the wave-tile kernel and the
// lane-tile kernel (bit-identical results) are timed once on the real factor and the fastest is kept; the losers' plans are released.
static void tri_autotune_wave(Handle &h) {
  if (h.n == 0 || !(h.wv.ready || h.lt.ready)) { h.tri_mode = 0; return; }
  cudaStream_t st = h.stream;
  const long long launch0 = h.st_launch, pcond0 = h.st_pcond;
  DBuf<double> a, b; a.ensure(h.n); b.ensure(h.n);
  B200_CUDA(cudaMemsetAsync(h.ctrl.p, 0, sizeof(Ctrl), st));
  B200_CUDA(cudaMemsetAsync(a.p, 0, (size_t)h.n * sizeof(double), st));
  float ms[3] = {0, 0, 0};
  const int modes[3] = {0, 3, 4};
  const bool have[3] = {true, h.wv.ready, h.lt.ready};
  int best = 0, napp = 0;
  for (int m = 0; m < 3; ++m) {
    if (!have[m]) continue;
    h.tri_mode = modes[m];
    lu_apply(h, b.p, a.p);
    B200_CUDA(cudaEventRecord(h.evf0, st));
    for (int r = 0; r < 3; ++r) lu_apply(h, b.p, a.p);
    B200_CUDA(cudaEventRecord(h.evf1, st));
    B200_CUDA(cudaStreamSynchronize(st));
    B200_CUDA(cudaEventElapsedTime(&ms[m], h.evf0, h.evf1));
    if (ms[m] < ms[best]) best = m;
    napp += 4;
  }
  B200_CUDA(cudaMemcpyAsync(h.h_ctrl, h.ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  B200_REQUIRE(h.h_ctrl->spin_timeout == 0, "triangular solve: dependency wait timed out while tuning");
  h.tri_mode = modes[best];
  if (h.tri_mode != 3) wave_release(h);
  if (h.tri_mode != 4) lane_release(h);
  if (getenv("B200_WAVE_DEBUG"))
    fprintf(stderr, "[tri] autotune: level kernel %.3f ms, wave tiles %.3f ms, lane tiles %.3f ms per application -> mode %d\n", ms[0] / 3, ms[1] / 3, ms[2] / 3, h.tri_mode);
  h.st_launch = launch0; h.st_pcond = pcond0;                      // tuning applications are not the caller's
  (void)napp;
  a.release(); b.release();
}

}  // namespace b200

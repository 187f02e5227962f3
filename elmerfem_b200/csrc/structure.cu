// Device-resident matrix structure: CRS mirror -> SELL-32 operand for SpMV, dependency levels and
// level-sorted SELL-32 plans for the ILU0 triangular solves.
//
// Reference data layout being replaced: Matrix_t (fem/src/Types.F90:193-283): Rows/Cols/Diag/Values,
// with ILURows/ILUCols/ILUDiag aliasing them for ILU(0) (fem/src/CRSMatrix.F90:3488-3491).
#include "common.cuh"
#include <thrust/scan.h>
#include <thrust/execution_policy.h>
#include <thrust/device_ptr.h>
#include <algorithm>
#include <climits>

namespace b200 {

// kind 0: whole row (SpMV operand)   start = rows[r],    len = rows[r+1]-rows[r]
// kind 1: strict lower part (L)      start = rows[r],    len = diag[r]-rows[r]
// kind 2: strict upper part (U)      start = diag[r]+1,  len = rows[r+1]-diag[r]-1
__global__ void k_slot_extent(int nslots, int n, const int *__restrict__ perm, const int *__restrict__ rows,
                              const int *__restrict__ diag, int kind, int *__restrict__ start, int *__restrict__ len) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nslots) return;
  int r = perm ? perm[slot] : (slot < n ? slot : -1);
  int s = 0, l = 0;
  if (r >= 0) {
    if (kind == 0) { s = rows[r]; l = rows[r + 1] - rows[r]; }
    else if (kind == 1) { s = rows[r]; l = diag[r] - rows[r]; }
    else { s = diag[r] + 1; l = rows[r + 1] - diag[r] - 1; }
  }
  start[slot] = s; len[slot] = l;
}

// one warp per slice: width = max len, stored as entries (width*32) for the scan
__global__ void k_slice_entries(int nslices, const int *__restrict__ len, long long *__restrict__ entries) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nslices) return;
  int l = len[w * 32 + lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l = max(l, __shfl_xor_sync(0xffffffffu, l, o));
  if (lane == 0) entries[w] = (long long)l * 32;
}

// Transposes one slice of CRS data into SELL order.  A warp walks its 32 rows entry by entry:
// lane owns one row; reads are served from L1 (a row is 1-3 cache lines), writes are coalesced.
template <class T, bool kSubBase>
__global__ void k_sell_fill(int nslots, const long long *__restrict__ ptr, const int *__restrict__ start,
                            const int *__restrict__ len, const T *__restrict__ src, T *__restrict__ dst) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nslots) return;
  int lane = threadIdx.x & 31;
  long long base = ptr[slot >> 5] + lane;
  int s = start[slot], l = len[slot];
  for (int j = 0; j < l; ++j) dst[base + (long long)j * 32] = src[s + j];
}

// pads: column = a valid index (own row / 0), value = 0; never read because loops stop at len.
template <class T>
__global__ void k_fill(long long n, T *p, T v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void k_remap_cols(long long n, int *cols, const int *__restrict__ map) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) cols[i] = map[cols[i]];
}

// S.start / S.len / S.perm already hold the per-slot extents; lays out the slices and copies the columns.
void sell_finish(Handle &h, Sell &S, int nslots, bool has_perm, const int *src_cols) {
  cudaStream_t st = h.stream;
  S.nslots = nslots; S.nslices = nslots / 32; S.has_perm = has_perm;
  S.ptr.ensure(S.nslices + 1);
  if (nslots == 0) { S.nstore = 0; B200_CUDA(cudaMemsetAsync(S.ptr.p, 0, sizeof(long long), st)); return; }
  k_slice_entries<<<(S.nslices * 32 + 255) / 256, 256, 0, st>>>(S.nslices, S.len.p, S.ptr.p);
  B200_CUDA(cudaMemsetAsync(S.ptr.p + S.nslices, 0, sizeof(long long), st));
  thrust::exclusive_scan(thrust::cuda::par.on(st), S.ptr.p, S.ptr.p + S.nslices + 1, S.ptr.p);
  B200_CUDA(cudaMemcpyAsync(&S.nstore, S.ptr.p + S.nslices, sizeof(long long), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  S.cols.ensure(S.nstore); S.vals.ensure(S.nstore);
  if (S.nstore) {
    B200_CUDA(cudaMemsetAsync(S.cols.p, 0, S.nstore * sizeof(int), st));
    B200_CUDA(cudaMemsetAsync(S.vals.p, 0, S.nstore * sizeof(double), st));
    k_sell_fill<int, false><<<(nslots + 255) / 256, 256, 0, st>>>(nslots, S.ptr.p, S.start.p, S.len.p, src_cols, S.cols.p);
  }
  B200_CUDA(cudaGetLastError());
}

static void build_sell(Handle &h, Sell &S, int kind, const int *d_perm, int nslots) {
  S.len.ensure(nslots); S.start.ensure(nslots);
  if (nslots) k_slot_extent<<<(nslots + 255) / 256, 256, 0, h.stream>>>(nslots, h.n, d_perm, h.d_rows.p, h.d_diag.p, kind, S.start.p, S.len.p);
  sell_finish(h, S, nslots, d_perm != nullptr, h.d_cols.p);
}

void sell_refresh_values(Handle &h, Sell &S, const double *crs_vals) {
  if (S.nslots == 0 || S.nstore == 0) return;
  k_sell_fill<double, false><<<(S.nslots + 255) / 256, 256, 0, h.stream>>>(S.nslots, S.ptr.p, S.start.p, S.len.p, crs_vals, S.vals.p);
  B200_CUDA(cudaGetLastError());
}

void structure_build(Handle &h) {
  int nslots = ((h.n + 31) / 32) * 32;
  build_sell(h, h.A, 0, nullptr, nslots);
}

// Dependency levels of the two triangular solves, computed on the host copy of the pattern (O(nnz),
// once per structure).  Forward level of row i = 1 + max level of the rows in its strict lower part
// (exactly the rows CRS_LUSolve's forward sweep, CRSMatrix.F90:4642-4649, and the IKJ elimination of
// CRS_IncompleteLU, 3624-3637, read before producing row i); backward likewise on the upper part
// (4653-4660).  Rows are then laid out level by level, each level padded to whole 32-row slices so
// that no warp ever holds two rows that depend on each other.
static void level_layout(int n, const std::vector<int> &level, int nlev, std::vector<int> &perm, int &nslots, bool descending,
                         std::vector<int> &gate, std::vector<int> &lvl_slices) {
  std::vector<long long> cnt(nlev + 1, 0);
  for (int i = 0; i < n; ++i) cnt[level[i] + 1]++;
  std::vector<long long> slice0(nlev + 1, 0);
  for (int l = 0; l < nlev; ++l) slice0[l + 1] = slice0[l] + (cnt[l + 1] + 31) / 32;
  long long ns = slice0[nlev] * 32;
  B200_REQUIRE(ns < 2147483647LL, "level layout exceeds int32 slots");
  nslots = (int)ns;
  perm.assign(nslots, -1);
  // per slice: its level; per level: number of slices (a slice of level l opens once level l - lookahead is finished)
  gate.assign(nslots / 32, 0);
  lvl_slices.assign(nlev, 0);
  for (int l = 0; l < nlev; ++l) {
    lvl_slices[l] = (int)(slice0[l + 1] - slice0[l]);
    for (long long sl = slice0[l]; sl < slice0[l + 1]; ++sl) gate[sl] = l;
  }
  std::vector<long long> fill(nlev, 0);
  if (!descending) {
    for (int i = 0; i < n; ++i) { int l = level[i]; perm[slice0[l] * 32 + fill[l]++] = i; }
  } else {
    for (int i = n - 1; i >= 0; --i) { int l = level[i]; perm[slice0[l] * 32 + fill[l]++] = i; }
  }
}

// Part layout of one sweep.  The rows are cut into P contiguous natural-order ranges ("parts", in the
// order the sweep visits them: ascending for the forward, descending for the backward sweep); a part
// only ever depends on itself and on parts visited earlier, so one CTA per part can walk its rows in
// level order while the parts form a pipeline.  Inside a part the rows are sorted by (global level,
// visiting order) and each level is padded to whole 32-row slices.  share[q] is the fraction of the
// work (matrix entries) given to part q: earlier parts start earlier and may take more.
static void part_layout(int n, const std::vector<int> &level, const std::vector<int> &rowlen, int P, double ramp, bool descending,
                        std::vector<int> &perm, int &nslots, std::vector<int> &part_slice_begin) {
  part_slice_begin.assign(P + 1, 0);
  perm.clear();
  if (n == 0) { nslots = 0; return; }
  // work prefix in visiting order
  std::vector<double> target(P + 1, 0.0);
  { double tot = 0; for (int q = 0; q < P; ++q) { double sh = 1.0 - (1.0 - ramp) * (P > 1 ? (double)q / (P - 1) : 0.0); tot += sh; target[q + 1] = tot; }
    for (int q = 0; q <= P; ++q) target[q] /= tot; }
  long long total = 0; for (int i = 0; i < n; ++i) total += rowlen[i] + 2;
  std::vector<int> cut(P + 1, 0);      // in visiting positions 0..n
  { long long acc = 0; int q = 1;
    for (int v = 0; v < n; ++v) {
      int i = descending ? n - 1 - v : v;
      acc += rowlen[i] + 2;
      while (q < P && (double)acc >= target[q] * (double)total) cut[q++] = v + 1;
    }
    while (q <= P) cut[q++] = n; }
  perm.reserve((size_t)n + (size_t)P * 64);
  std::vector<int> cnt, pos;
  for (int q = 0; q < P; ++q) {
    part_slice_begin[q] = (int)(perm.size() / 32);
    const int v0 = cut[q], v1 = cut[q + 1];
    if (v1 <= v0) continue;
    int lmin = INT_MAX, lmax = -1;
    for (int v = v0; v < v1; ++v) { int i = descending ? n - 1 - v : v; lmin = std::min(lmin, level[i]); lmax = std::max(lmax, level[i]); }
    cnt.assign((size_t)(lmax - lmin + 2), 0);
    for (int v = v0; v < v1; ++v) { int i = descending ? n - 1 - v : v; cnt[level[i] - lmin + 1]++; }
    // padded start of each level
    pos.assign((size_t)(lmax - lmin + 1), 0);
    size_t base = perm.size(), off = 0;
    for (int l = 0; l <= lmax - lmin; ++l) { pos[l] = (int)off; off += ((size_t)cnt[l + 1] + 31) / 32 * 32; }
    B200_REQUIRE(base + off < 2147483647ULL, "part layout exceeds int32 slots");
    perm.resize(base + off, -1);
    for (int v = v0; v < v1; ++v) { int i = descending ? n - 1 - v : v; perm[base + pos[level[i] - lmin]++] = i; }
  }
  part_slice_begin[P] = (int)(perm.size() / 32);
  nslots = (int)perm.size();
}

// Column ids of a part-mode plan -> source codes, plus the per-slice dataflow metadata of k_sptrsv_tok.
//   code >= 0 : slot in the global (L2-resident) solve vector
//   code <  0 : ~(ring index | same_round << 16): the producer belongs to the same part and ran in the
//               same round (needs the mbarrier token) or at most TRI_RR-2 rounds earlier (final)
// meta[slice].x = pred | succ << 16: warps of the same round this slice waits for / has to wake;
// meta[slice].y = npre: entries every lane may consume before the token (none of them is same-round).
__global__ void k_tri_codes(int nslots, const long long *__restrict__ ptr, const int *__restrict__ len, int *cols,
                            const int *__restrict__ slot_of_row, const int *__restrict__ part_of_slice,
                            const int *__restrict__ part_begin, int2 *meta) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nslots) return;                                  // nslots is a multiple of 32: whole warps leave
  const int lane = threadIdx.x & 31, slice = slot >> 5;
  const long long base = ptr[slice] + lane;
  const int l = len[slot];
  const int part = part_of_slice[slice], b0 = part_begin[part];
  const int kc = slice - b0, rc = kc / TRI_NW;
  int first_same = 255; unsigned predbits = 0;
  for (int j = 0; j < l; ++j) {
    const int sp = slot_of_row[cols[base + (long long)j * 32]];
    const int slp = sp >> 5;
    int code = sp;
    if (part_of_slice[slp] == part) {
      const int kp = slp - b0, rp = kp / TRI_NW;
      if (rc - rp <= TRI_RR - 2) {
        const int same = (rp == rc) ? 1 : 0;
        code = ~(((kp % (TRI_RR * TRI_NW)) * 32 + (sp & 31)) | (same << 16));
        if (same) {
          first_same = min(first_same, j);
          predbits |= 1u << (kp % TRI_NW);
          atomicOr(&meta[slp].x, (int)((1u << (kc % TRI_NW)) << 16));
        }
      }
    }
    cols[base + (long long)j * 32] = code;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    first_same = min(first_same, __shfl_xor_sync(0xffffffffu, first_same, o));
    predbits |= __shfl_xor_sync(0xffffffffu, predbits, o);
  }
  if (lane == 0) { if (predbits) atomicOr(&meta[slice].x, (int)predbits); meta[slice].y = first_same; }
}

void tri_analyse(Handle &h) {
  if (h.tri_ready) return;
  const int n = h.n;
  const std::vector<int> &rows = h.h_rows, &cols = h.h_cols, &diag = h.h_diag;
  std::vector<int> lf(n, 0), lb(n, 0);
  int nlf = 0, nlb = 0;
  for (int i = 0; i < n; ++i) {
    int l = 0;
    for (int p = rows[i]; p < diag[i]; ++p) l = std::max(l, lf[cols[p]] + 1);
    lf[i] = l; nlf = std::max(nlf, l + 1);
  }
  for (int i = n - 1; i >= 0; --i) {
    int l = 0;
    for (int p = diag[i] + 1; p < rows[i + 1]; ++p) l = std::max(l, lb[cols[p]] + 1);
    lb[i] = l; nlb = std::max(nlb, l + 1);
  }
  if (n == 0) { nlf = nlb = 0; }
  h.nlev_f = nlf; h.nlev_b = nlb;
  std::vector<int> pf, pb; int nsf = 0, nsb = 0;
  std::vector<int> gf, gb, cf, cb;
  // global level order: the order in which the factorisation visits the rows, and the plan of the level-mode solves
  level_layout(n, lf, nlf, pf, nsf, false, gf, cf);
  h.d_order_f.ensure(nsf); h.n_order_f = nsf;
  if (nsf) B200_CUDA(cudaMemcpyAsync(h.d_order_f.p, pf.data(), (size_t)nsf * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaStreamSynchronize(h.stream));
  std::vector<int> pbf, pbb;           // part boundaries in slices
  if (h.tri_mode == 1) {
    int maxp = h.tri_max_parts > 0 ? h.tri_max_parts : NUM_SMS;
    int P = std::max(1, std::min(maxp, (n + h.tri_rows_per_part_min - 1) / std::max(1, h.tri_rows_per_part_min)));
    std::vector<int> ll(n), lu(n);
    for (int i = 0; i < n; ++i) { ll[i] = diag[i] - rows[i]; lu[i] = rows[i + 1] - diag[i] - 1; }
    part_layout(n, lf, ll, P, h.tri_ramp, false, pf, nsf, pbf);
    part_layout(n, lb, lu, P, h.tri_ramp, true, pb, nsb, pbb);
    h.tri_parts_f = h.tri_parts_b = P;
    h.d_part_begin_f.ensure(P + 1); h.d_part_begin_b.ensure(P + 1);
    B200_CUDA(cudaMemcpyAsync(h.d_part_begin_f.p, pbf.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice, h.stream));
    B200_CUDA(cudaMemcpyAsync(h.d_part_begin_b.p, pbb.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  } else {
    level_layout(n, lb, nlb, pb, nsb, true, gb, cb);
    h.L.gate.ensure(gf.size()); h.U.gate.ensure(gb.size());
    h.d_lvlcnt_f.ensure(cf.size()); h.d_lvlcnt_b.ensure(cb.size());
    if (!cf.empty()) B200_CUDA(cudaMemcpyAsync(h.d_lvlcnt_f.p, cf.data(), cf.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
    if (!cb.empty()) B200_CUDA(cudaMemcpyAsync(h.d_lvlcnt_b.p, cb.data(), cb.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
    if (!gf.empty()) B200_CUDA(cudaMemcpyAsync(h.L.gate.p, gf.data(), gf.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
    if (!gb.empty()) B200_CUDA(cudaMemcpyAsync(h.U.gate.p, gb.data(), gb.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  }
  h.L.perm.ensure(nsf); h.U.perm.ensure(nsb);
  h.tri_counters.ensure(((size_t)nlf + nlb + 2) * 32);
  if (nsf) B200_CUDA(cudaMemcpyAsync(h.L.perm.p, pf.data(), (size_t)nsf * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  if (nsb) B200_CUDA(cudaMemcpyAsync(h.U.perm.p, pb.data(), (size_t)nsb * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaStreamSynchronize(h.stream));
  build_sell(h, h.L, 1, h.L.perm.p, nsf);
  build_sell(h, h.U, 2, h.U.perm.p, nsb);
  // The solves keep their vectors in slot order: a warp's 32 rows store one coalesced line, and the
  // entries it gathers sit in neighbouring slots instead of being strewn over the natural numbering.
  // Columns of L/U are renumbered to slots (level mode) or to source codes (part mode); the backward
  // sweep reads its right-hand side (the forward result, L-slot order) through urhs.
  {
    std::vector<int> slotL(std::max(n, 1), 0), slotU(std::max(n, 1), 0), urhs(std::max(nsb, 1), 0);
    for (int s = 0; s < nsf; ++s) if (pf[s] >= 0) slotL[pf[s]] = s;
    for (int s = 0; s < nsb; ++s) if (pb[s] >= 0) { slotU[pb[s]] = s; urhs[s] = slotL[pb[s]]; }
    DBuf<int> dL, dU; dL.ensure(n); dU.ensure(n); h.d_urhs.ensure(nsb);
    if (n) {
      B200_CUDA(cudaMemcpyAsync(dL.p, slotL.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h.stream));
      B200_CUDA(cudaMemcpyAsync(dU.p, slotU.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h.stream));
      B200_CUDA(cudaMemcpyAsync(h.d_urhs.p, urhs.data(), (size_t)nsb * sizeof(int), cudaMemcpyHostToDevice, h.stream));
      if (h.tri_mode == 1) {
        auto part_of = [&](const std::vector<int> &pbeg, int nslices) {
          std::vector<int> v((size_t)std::max(nslices, 1), 0);
          for (int q = 0; q + 1 < (int)pbeg.size(); ++q) for (int sl = pbeg[q]; sl < pbeg[q + 1]; ++sl) v[sl] = q;
          return v;
        };
        std::vector<int> posf = part_of(pbf, nsf / 32), posb = part_of(pbb, nsb / 32);
        DBuf<int> dpf, dpb; dpf.ensure(posf.size()); dpb.ensure(posb.size());
        B200_CUDA(cudaMemcpyAsync(dpf.p, posf.data(), posf.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
        B200_CUDA(cudaMemcpyAsync(dpb.p, posb.data(), posb.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
        h.d_meta_f.ensure((size_t)std::max(nsf / 32, 1) * 2); h.d_meta_b.ensure((size_t)std::max(nsb / 32, 1) * 2);
        B200_CUDA(cudaMemsetAsync(h.d_meta_f.p, 0, (size_t)std::max(nsf / 32, 1) * 2 * sizeof(int), h.stream));
        B200_CUDA(cudaMemsetAsync(h.d_meta_b.p, 0, (size_t)std::max(nsb / 32, 1) * 2 * sizeof(int), h.stream));
        if (nsf) k_tri_codes<<<(nsf + 255) / 256, 256, 0, h.stream>>>(nsf, h.L.ptr.p, h.L.len.p, h.L.cols.p, dL.p, dpf.p, h.d_part_begin_f.p, (int2 *)h.d_meta_f.p);
        if (nsb) k_tri_codes<<<(nsb + 255) / 256, 256, 0, h.stream>>>(nsb, h.U.ptr.p, h.U.len.p, h.U.cols.p, dU.p, dpb.p, h.d_part_begin_b.p, (int2 *)h.d_meta_b.p);
        {                                                 // L slot -> U slot of the same row: the forward sweep scatters its result there
          std::vector<int> l2u((size_t)std::max(nsf, 1), 0);
          for (int sl = 0; sl < nsf; ++sl) if (pf[sl] >= 0) l2u[sl] = slotU[pf[sl]];
          h.d_l2u.ensure(l2u.size());
          B200_CUDA(cudaMemcpyAsync(h.d_l2u.p, l2u.data(), l2u.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
          B200_CUDA(cudaStreamSynchronize(h.stream));
        }
        h.d_bl.ensure(std::max(nsf, 1)); h.d_bu.ensure(std::max(nsb, 1));
        B200_CUDA(cudaMemsetAsync(h.d_bu.p, 0, (size_t)std::max(nsb, 1) * sizeof(double), h.stream));
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaStreamSynchronize(h.stream));
        dpf.release(); dpb.release();
      } else {
        if (h.L.nstore) k_remap_cols<<<NUM_SMS * 8, 256, 0, h.stream>>>(h.L.nstore, h.L.cols.p, dL.p);
        if (h.U.nstore) k_remap_cols<<<NUM_SMS * 8, 256, 0, h.stream>>>(h.U.nstore, h.U.cols.p, dU.p);
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaStreamSynchronize(h.stream));
      }
    }
    dL.release(); dU.release();
  }
  h.d_yl.ensure(nsf); h.d_xu.ensure(nsb);
  h.d_dinv_slot.ensure(nsb);
  {
    int mw = 0;
    for (int i = 0; i < n; ++i) mw = std::max(mw, std::max(diag[i] - rows[i], rows[i + 1] - diag[i] - 1));
    h.tri_maxw = mw;
  }
  h.d_rowdone.ensure(n);
  h.h_level_f.swap(lf);
  h.tri_ready = true;
}

}  // namespace b200

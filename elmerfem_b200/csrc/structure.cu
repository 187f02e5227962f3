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
// ralign: rows are pushed to the END of their slice (entry j of a row of length l at position W - l + j),
// so that the last entries of all 32 rows sit in the same positions (forward triangular plan).
template <class T, bool kSubBase>
__global__ void k_sell_fill(int nslots, const long long *__restrict__ ptr, const int *__restrict__ start,
                            const int *__restrict__ len, const T *__restrict__ src, T *__restrict__ dst, bool ralign) {
  int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nslots) return;
  int lane = threadIdx.x & 31;
  long long base = ptr[slot >> 5] + lane;
  int s = start[slot], l = len[slot];
  if (ralign) base += (long long)((int)((ptr[(slot >> 5) + 1] - ptr[slot >> 5]) >> 5) - l) * 32;
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
    k_sell_fill<int, false><<<(nslots + 255) / 256, 256, 0, st>>>(nslots, S.ptr.p, S.start.p, S.len.p, src_cols, S.cols.p, S.ralign);
  }
  B200_CUDA(cudaGetLastError());
}

static void build_sell(Handle &h, Sell &S, int kind, const int *d_perm, int nslots) {
  S.len.ensure(nslots); S.start.ensure(nslots);
  const int *rows = kind == 0 ? h.d_rows.p : h.d_lrows(), *diag = kind == 0 ? h.d_diag.p : h.d_ldiag();   // L/U plans live on the ILU pattern
  if (nslots) k_slot_extent<<<(nslots + 255) / 256, 256, 0, h.stream>>>(nslots, h.n, d_perm, rows, diag, kind, S.start.p, S.len.p);
  sell_finish(h, S, nslots, d_perm != nullptr, kind == 0 ? h.d_cols.p : h.d_lcols());
}

void sell_refresh_values(Handle &h, Sell &S, const double *crs_vals) {
  if (S.nslots == 0 || S.nstore == 0) return;
  k_sell_fill<double, false><<<(S.nslots + 255) / 256, 256, 0, h.stream>>>(S.nslots, S.ptr.p, S.start.p, S.len.p, crs_vals, S.vals.p, S.ralign);
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

// Node-lane layout for matrices with ND interleaved dofs per node (h.ndeg > 1; rows of node I are I*ND .. I*ND+ND-1, and every row holds
// its node's full diagonal block): levels are those of the NODE graph, and the ND rows of a node sit in ADJACENT LANES of one slice
// (lane = ni * ND + d, 32 / ND nodes per slice).  k_sptrsv_wide_node then passes a node's own results between lanes by shuffle, so the
// dependency chain through L2 has one hop per NODE level instead of one per row level (3784 -> 946 on the 4-dof cavity operand,
// 2904 -> 968 on the 3-dof beam).  Slices of this layout hold mutually dependent rows: only the node-aware sweeps may run on it (the
// row-per-thread polling kernels would wait for a value of their own warp); the factorisation (warp per row, slot order = a topological
// order) is unaffected.
static void node_lane_layout(int nnodes, int ND, const std::vector<int> &nlevel, int nlev, std::vector<int> &perm, int &nslots, bool backward,
                             std::vector<int> &gate, std::vector<int> &lvl_slices) {
  const int NPW = 32 / ND;
  std::vector<long long> cnt(nlev + 1, 0);
  for (int I = 0; I < nnodes; ++I) cnt[nlevel[I] + 1]++;
  std::vector<long long> slice0(nlev + 1, 0);
  for (int l = 0; l < nlev; ++l) slice0[l + 1] = slice0[l] + (cnt[l + 1] + NPW - 1) / NPW;
  const long long ns = slice0[nlev] * 32;
  B200_REQUIRE(ns < 2147483647LL, "node layout exceeds int32 slots");
  nslots = (int)ns;
  perm.assign(nslots, -1);
  gate.assign(nslots / 32, 0);
  lvl_slices.assign(nlev, 0);
  for (int l = 0; l < nlev; ++l) {
    lvl_slices[l] = (int)(slice0[l + 1] - slice0[l]);
    for (long long sl = slice0[l]; sl < slice0[l + 1]; ++sl) gate[sl] = l;
  }
  std::vector<long long> fill(nlev, 0);
  auto place = [&](int I) {
    const int l = nlevel[I];
    const long long k = fill[l]++;
    const long long sl = slice0[l] + k / NPW; const int ni = (int)(k % NPW);
    for (int d = 0; d < ND; ++d) perm[sl * 32 + ni * ND + d] = I * ND + d;
  };
  if (!backward) for (int I = 0; I < nnodes; ++I) place(I);
  else for (int I = nnodes - 1; I >= 0; --I) place(I);
}

// One round of InitializeILU1 (CRSMatrix.F90:3664-3795) on a 0-based pattern: row i keeps its entries and gains the
// columns of the upper parts of the rows k < i it held BEFORE the round (fills of the round do not cascade);
// columns ascending.
static void ilu1_round(int n, const std::vector<int> &rows, const std::vector<int> &cols, const std::vector<int> &diag,
                       std::vector<int> &r2, std::vector<int> &c2, std::vector<int> &d2) {
  std::vector<unsigned char> C(n, 0);
  r2.assign((size_t)n + 1, 0); d2.assign(n, 0); c2.clear(); c2.reserve(cols.size() * 2);
  std::vector<int> added;
  for (int i = 0; i < n; ++i) {
    for (int k = rows[i]; k < rows[i + 1]; ++k) C[cols[k]] = 1;
    added.clear();
    for (int m = rows[i]; m < diag[i]; ++m) {              // the rows k < i of the pattern (flag 1), ascending
      const int k = cols[m];
      for (int l = diag[k] + 1; l < rows[k + 1]; ++l) { const int j = cols[l]; if (C[j] == 0) { C[j] = 2; added.push_back(j); } }
    }
    std::sort(added.begin(), added.end());
    size_t a = 0;
    for (int k = rows[i]; k < rows[i + 1] || a < added.size();) {   // merge, ascending
      int c;
      if (a < added.size() && (k >= rows[i + 1] || added[a] < cols[k])) c = added[a++]; else c = cols[k++];
      if (c == i) d2[i] = (int)c2.size();
      c2.push_back(c); C[c] = 0;
    }
    B200_REQUIRE(c2.size() < 2147483647ULL, "ILU(n) pattern exceeds int32 entries");
    r2[i + 1] = (int)c2.size();
  }
}

void ilu_pattern_build(Handle &h) {
  if (!h.ilu_sep() || h.ilu_pat_ready) return;
  B200_REQUIRE(!h.ilut, "ILUT: the pattern is installed by the factorisation");
  const int n = h.n;
  std::vector<int> r = h.h_rows, c = h.h_cols, d = h.h_diag, r2, c2, d2;
  if (h.bilu_blocks > 1) {                                 // CRS_BlockDiagonal: keep the entries with MOD(i,Blocks) == MOD(j,Blocks)
    const int B = h.bilu_blocks;
    r2.assign((size_t)n + 1, 0); c2.clear(); d2.assign(n, 0);
    for (int i = 0; i < n; ++i) {
      for (int p = r[i]; p < r[i + 1]; ++p) if (i % B == c[p] % B) { if (c[p] == i) d2[i] = (int)c2.size(); c2.push_back(c[p]); }
      r2[i + 1] = (int)c2.size();
    }
    r.swap(r2); c.swap(c2); d.swap(d2);
  }
  for (int round = 0; round < h.ilu_order; ++round) { ilu1_round(n, r, c, d, r2, c2, d2); r.swap(r2); c.swap(c2); d.swap(d2); }
  std::vector<int> src(c.size(), -1);
#pragma omp parallel for
  for (int i = 0; i < n; ++i) {
    int p = h.h_rows[i];
    for (int q = r[i]; q < r[i + 1]; ++q) {
      while (p < h.h_rows[i + 1] && h.h_cols[p] < c[q]) ++p;
      if (p < h.h_rows[i + 1] && h.h_cols[p] == c[q]) src[q] = p;
    }
  }
  h.ilu_nnz = (long long)c.size();
  h.dl_rows.ensure((size_t)n + 1); h.dl_cols.ensure(c.size()); h.dl_diag.ensure(std::max(n, 1)); h.dl_src.ensure(c.size());
  B200_CUDA(cudaMemcpyAsync(h.dl_rows.p, r.data(), ((size_t)n + 1) * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  if (!c.empty()) {
    B200_CUDA(cudaMemcpyAsync(h.dl_cols.p, c.data(), c.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
    B200_CUDA(cudaMemcpyAsync(h.dl_src.p, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  }
  if (n) B200_CUDA(cudaMemcpyAsync(h.dl_diag.p, d.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaStreamSynchronize(h.stream));
  h.hl_rows.swap(r); h.hl_cols.swap(c); h.hl_diag.swap(d);
  h.ilu_pat_ready = true;
}

void ilu_invalidate(Handle &h) {
  h.ilu_valid = h.ilu_exists = false; h.tri_ready = false; h.ilu_pat_ready = false;
  h.grid_ilu = h.grid_tri_l = h.grid_tri_u = 0; h.grid_ilu_kern = nullptr; h.ilu_map_maxu = 0; h.ilu_map_tried = false; h.ilu_reg_ok = -1; h.d_ilu_pos.release(); h.d_ilu_posptr.release();
  wave_release(h);
  lane_release(h);
  ichol_release(h);
  h.tri_mode = h.tri_mode_cfg;
}

// Incomplete Cholesky solve (CRSMatrix.F90:4618-4638).  The backward loop is column-oriented: row i, taken from n down to 1, scales
// b(i) and subtracts L(i, c) b(i) from every b(c) of its lower part.  Unknown c therefore receives its updates from the rows i > c that
// hold column c, in DESCENDING i, before it is scaled.  ch_ptr / ch_row / ch_pos list exactly that per column (the gather form of the
// same sums, same order); the dependency levels of the sweep follow from the lists.  Needs the row-level forward plan.
// The tile kernels' entry streams are refilled after every factorisation.  The row-wise scatter (k_wave_fill / k_lane_fill: 27 scattered
// 8-byte stores per row) took 11 ms on the 200^3 problem; here the same scatter runs ONCE per structure on an array of indices (entry q
// holds q + 1), which turns into a map stream entry -> ILU position, and every refill is a coalesced gather through that map.
__global__ void k_stream_iota1(long long n, double *__restrict__ a) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = (double)(i + 1);
}
__global__ void k_stream_map(long long n, const double *__restrict__ S, int *__restrict__ map) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) map[i] = (int)S[i] - 1;
}
__global__ void k_stream_gather(long long n, const int *__restrict__ map, const double *__restrict__ ilu, double *__restrict__ S) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int m = map[i];
    S[i] = m >= 0 ? ilu[m] : 0.0;
  }
}
void stream_iota1(Handle &h, long long n, double *a) { if (n) k_stream_iota1<<<NUM_SMS * 8, 256, 0, h.stream>>>(n, a); B200_CUDA(cudaGetLastError()); }
void stream_map_build(Handle &h, long long n, const double *S, int *map) { if (n) k_stream_map<<<NUM_SMS * 8, 256, 0, h.stream>>>(n, S, map); B200_CUDA(cudaGetLastError()); }
void stream_gather(Handle &h, long long n, const int *map, const double *ilu, double *S) { if (n) k_stream_gather<<<NUM_SMS * 8, 256, 0, h.stream>>>(n, map, ilu, S); B200_CUDA(cudaGetLastError()); }

void ichol_release(Handle &h) {
  h.ch_ptr.release(); h.ch_row.release(); h.ch_pos.release(); h.ch_perm.release(); h.ch_gate.release(); h.ch_lvlcnt.release(); h.ch_counters.release();
  h.ch_y.release(); h.ch_x.release(); h.ch_ready = false;
}
void ichol_analyse(Handle &h) {
  if (h.ch_ready) return;
  if (h.tri_node) { h.tri_node_off = true; ilu_invalidate(h); }     // the sweeps poll per row: row-level layout only
  tri_analyse(h);
  const int n = h.n;
  const std::vector<int> &rows = h.lrows(), &cols = h.lcols(), &diag = h.ldiag();
  std::vector<int> ptr((size_t)n + 1, 0);
  for (int i = 0; i < n; ++i) for (int p = rows[i]; p < diag[i]; ++p) ptr[cols[p] + 1]++;
  for (int c = 0; c < n; ++c) ptr[c + 1] += ptr[c];
  std::vector<int> row(std::max(ptr[n], 1)), pos(std::max(ptr[n], 1)), fill(ptr.begin(), ptr.end() - 1);
  for (int i = n - 1; i >= 0; --i) for (int p = rows[i]; p < diag[i]; ++p) { const int q = fill[cols[p]]++; row[q] = i; pos[q] = p; }   // rows descending
  std::vector<int> lv(n, 0); int nlev = 0;
  for (int c = n - 1; c >= 0; --c) {
    int l = 0;
    for (int q = ptr[c]; q < ptr[c + 1]; ++q) l = std::max(l, lv[row[q]] + 1);
    lv[c] = l; nlev = std::max(nlev, l + 1);
  }
  if (n == 0) nlev = 0;
  std::vector<int> perm, gate, cnt; int nslots = 0;
  level_layout(n, lv, nlev, perm, nslots, true, gate, cnt);
  h.ch_nlev = nlev; h.ch_nslices = nslots / 32;
  h.ch_ptr.ensure(ptr.size()); h.ch_row.ensure(row.size()); h.ch_pos.ensure(pos.size());
  h.ch_perm.ensure(std::max(nslots, 1)); h.ch_gate.ensure(std::max<size_t>(gate.size(), 1)); h.ch_lvlcnt.ensure(std::max<size_t>(cnt.size(), 1));
  h.ch_counters.ensure(((size_t)h.nlev_f + nlev + 2) * 32);
  h.ch_y.ensure(n); h.ch_x.ensure(n);
  cudaStream_t st = h.stream;
  B200_CUDA(cudaMemcpyAsync(h.ch_ptr.p, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  B200_CUDA(cudaMemcpyAsync(h.ch_row.p, row.data(), row.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  B200_CUDA(cudaMemcpyAsync(h.ch_pos.p, pos.data(), pos.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  if (nslots) B200_CUDA(cudaMemcpyAsync(h.ch_perm.p, perm.data(), (size_t)nslots * sizeof(int), cudaMemcpyHostToDevice, st));
  if (!gate.empty()) B200_CUDA(cudaMemcpyAsync(h.ch_gate.p, gate.data(), gate.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  if (!cnt.empty()) B200_CUDA(cudaMemcpyAsync(h.ch_lvlcnt.p, cnt.data(), cnt.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  B200_CUDA(cudaStreamSynchronize(st));
  h.ch_ready = true;
}

void tri_analyse(Handle &h) {
  if (h.tri_ready) return;
  ilu_pattern_build(h);
  const int n = h.n;
  const std::vector<int> &rows = h.lrows(), &cols = h.lcols(), &diag = h.ldiag();
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
  std::vector<int> pf, pb; int nsf = 0, nsb = 0;
  std::vector<int> gf, gb, cf, cb;
  // Node-lane plans are used where they were measured to win: rows wider than 48 entries per triangle (4-dof Navier-Stokes operands:
  // 17.3 -> 8.0 ms per ILU0 application on the 96^3 cavity), for which the row-level path is the chunked generic loop.  On the 3-dof beam
  // (rows of 40: register-resident row-level kernel) forward wins (4.1 -> 2.7 ms) but backward loses (4.1 -> 7.3 ms) and a mixed plan is
  // no faster overall (8.5 vs 8.2 ms): the row-level layout stays.  B200_TRI_NODE = 0 never, 2 whenever the structure allows.
  const int node_cfg = getenv("B200_TRI_NODE") ? atoi(getenv("B200_TRI_NODE")) : 1;
  const bool node_ok = node_cfg != 0;
  const int ND = h.ndeg;
  int maxw = 0;
  for (int i = 0; i < n; ++i) maxw = std::max(maxw, std::max(diag[i] - rows[i], rows[i + 1] - diag[i] - 1));
  h.tri_node = 0;
  bool node = node_ok && !h.tri_node_off && ND >= 2 && ND <= 6 && n > 0 && n % ND == 0 && maxw <= 64 && (maxw > 48 || node_cfg == 2);
  if (node) {                                                     // every row must hold its node's full diagonal block, adjacent to the diagonal
#pragma omp parallel for schedule(static) reduction(&& : node)
    for (int i = 0; i < n; ++i) {
      const int I = i / ND, d = i - I * ND;
      bool ok = (diag[i] - rows[i] >= d) && (rows[i + 1] - diag[i] - 1 >= ND - 1 - d);
      for (int t = 0; ok && t < d; ++t) ok = cols[diag[i] - d + t] == I * ND + t;
      for (int t = 0; ok && t < ND - 1 - d; ++t) ok = cols[diag[i] + 1 + t] == i + 1 + t;
      node = node && ok;
    }
  }
  if (node) {
    const int nn = n / ND;
    std::vector<int> nf(nn, 0), nb(nn, 0);
    int nnf = 0, nnb = 0;
    for (int I = 0; I < nn; ++I) {
      int l = 0;
      for (int i = I * ND; i < I * ND + ND; ++i)
        for (int p = rows[i]; p < diag[i]; ++p) { const int J = cols[p] / ND; if (J != I) l = std::max(l, nf[J] + 1); }
      nf[I] = l; nnf = std::max(nnf, l + 1);
    }
    for (int I = nn - 1; I >= 0; --I) {
      int l = 0;
      for (int i = I * ND; i < I * ND + ND; ++i)
        for (int p = diag[i] + 1; p < rows[i + 1]; ++p) { const int J = cols[p] / ND; if (J != I) l = std::max(l, nb[J] + 1); }
      nb[I] = l; nnb = std::max(nnb, l + 1);
    }
    // (B200_TRI_NODE_U = 0: backward plan in the row-level layout, for experiments)
    bool node_u = true;
    if (getenv("B200_TRI_NODE_U")) node_u = atoi(getenv("B200_TRI_NODE_U")) != 0;
    node_lane_layout(nn, ND, nf, nnf, pf, nsf, false, gf, cf);
    nlf = nnf;
    if (node_u) { node_lane_layout(nn, ND, nb, nnb, pb, nsb, true, gb, cb); nlb = nnb; }
    else level_layout(n, lb, nlb, pb, nsb, true, gb, cb);
    h.tri_node = ND; h.tri_node_u = node_u;
  } else {
    level_layout(n, lf, nlf, pf, nsf, false, gf, cf);
    level_layout(n, lb, nlb, pb, nsb, true, gb, cb);
  }
  h.nlev_f = nlf; h.nlev_b = nlb;
  h.L.perm.ensure(nsf); h.U.perm.ensure(nsb);
  h.L.gate.ensure(gf.size()); h.U.gate.ensure(gb.size());
  h.d_lvlcnt_f.ensure(cf.size()); h.d_lvlcnt_b.ensure(cb.size());
  h.tri_counters.ensure(((size_t)nlf + nlb + 2) * 32);
  if (!cf.empty()) B200_CUDA(cudaMemcpyAsync(h.d_lvlcnt_f.p, cf.data(), cf.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  if (!cb.empty()) B200_CUDA(cudaMemcpyAsync(h.d_lvlcnt_b.p, cb.data(), cb.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  if (!gf.empty()) B200_CUDA(cudaMemcpyAsync(h.L.gate.p, gf.data(), gf.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  if (!gb.empty()) B200_CUDA(cudaMemcpyAsync(h.U.gate.p, gb.data(), gb.size() * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  if (nsf) B200_CUDA(cudaMemcpyAsync(h.L.perm.p, pf.data(), (size_t)nsf * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  if (nsb) B200_CUDA(cudaMemcpyAsync(h.U.perm.p, pb.data(), (size_t)nsb * sizeof(int), cudaMemcpyHostToDevice, h.stream));
  B200_CUDA(cudaStreamSynchronize(h.stream));
  h.L.ralign = true;
  build_sell(h, h.L, 1, h.L.perm.p, nsf);
  build_sell(h, h.U, 2, h.U.perm.p, nsb);
  // The solves keep their vectors in level (slot) order: a warp's 32 rows store one coalesced line, and
  // the entries it gathers from the previous levels sit in neighbouring slots instead of being strewn
  // over the natural numbering.  Columns of L/U are renumbered to slots; the backward sweep reads its
  // right-hand side (the forward result, L-slot order) through urhs.
  {
    std::vector<int> slotL(std::max(n, 1), 0), slotU(std::max(n, 1), 0), urhs(std::max(nsb, 1), 0);
    for (int s = 0; s < nsf; ++s) if (pf[s] >= 0) slotL[pf[s]] = s;
    for (int s = 0; s < nsb; ++s) if (pb[s] >= 0) { slotU[pb[s]] = s; urhs[s] = slotL[pb[s]]; }
    DBuf<int> dL, dU; dL.ensure(n); dU.ensure(n); h.d_urhs.ensure(nsb);
    if (n) {
      B200_CUDA(cudaMemcpyAsync(dL.p, slotL.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h.stream));
      B200_CUDA(cudaMemcpyAsync(dU.p, slotU.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h.stream));
      B200_CUDA(cudaMemcpyAsync(h.d_urhs.p, urhs.data(), (size_t)nsb * sizeof(int), cudaMemcpyHostToDevice, h.stream));
      if (h.L.nstore) k_remap_cols<<<NUM_SMS * 8, 256, 0, h.stream>>>(h.L.nstore, h.L.cols.p, dL.p);
      if (h.U.nstore) k_remap_cols<<<NUM_SMS * 8, 256, 0, h.stream>>>(h.U.nstore, h.U.cols.p, dU.p);
      B200_CUDA(cudaGetLastError());
      B200_CUDA(cudaStreamSynchronize(h.stream));
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

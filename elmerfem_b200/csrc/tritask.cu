// Task-mode triangular solves for the ILU(0) preconditioner: CRS_LUSolve, fem/src/CRSMatrix.F90:4590-4663.
//
// The level-mode kernel (precond.cu) pays one L2 hand-off per dependency level: 2 x 1401 hops of
// ~0.95 us on the 200^3 heat problem.  Here the rows are cut into TASKS = contiguous ranges of the
// sweep order (natural order for L, reversed for U).  One warp owns a task and walks through it in
// STEPS: a step is up to 32 rows of the task that are independent of each other once the previous
// steps are done (local dependency levels of the range).  Dependencies
//   * inside the task, at most TT_D-1 steps back: read from a per-warp shared-memory ring (no L2 hop);
//   * anywhere else: read from the result vector in L2, which is pre-filled with a NaN sentinel;
//     the values for step k+1 are requested while step k computes, and re-polled only if still unset.
// Because every dependency of a range points into a lower range, the task graph is acyclic: a task
// falls behind its producers by one L2 round trip ONCE (not once per level) and then streams at the
// pace of its own arithmetic.  Warp w runs tasks w, w+NW, ... in increasing order on a co-resident
// (cooperative) grid: the lowest unfinished task is always running and only waits on finished rows,
// so the scheme cannot deadlock for any matrix.
//
// Every row still performs the reference's operations in the reference's order (entries left to
// right, separate multiply / subtract roundings, inverse diagonal applied last): results are
// bit-identical to the level-mode kernel and to the CPU loop.
//
// Matrix entries travel as one byte stream per sweep, cut into per-step blocks
//   [header 16 B: nr, W, E, -][vals E x f64][codes E x u32][dinv nr4 x f64 (U only)][row ids nr4 x i32]
// entry j of the row on lane l at index j*nr + l.  Blocks are copied to shared memory with cp.async
// TT_P1 steps ahead of their use.  code = 0x80000000 | age << 8 | lane  (ring), 0xFFFFFFFF (padding)
// or the natural index of the row to read from the result vector.
#include "common.cuh"
#include "kernels.cuh"
#include <algorithm>
#include <omp.h>

namespace b200 {

constexpr unsigned TT_LOCAL = 0x80000000u;
constexpr unsigned TT_PAD = 0xFFFFFFFFu;
constexpr long long TT_SPIN_LIMIT = 1LL << 26;

static inline int ru4(int x) { return (x + 3) & ~3; }

// -----------------------------------------------------------------------------------------------
// host: plan construction
struct TaskSizes { long long steps = 0, bytes = 0, fills = 0; int maxblock = 0, maxw = 0; };
struct TaskScratch { std::vector<int> lev, cnt, stepb, order, slot; };

// Builds (or only measures, EMIT = false) the steps of the task holding sweep positions [q0, q1).
template <bool EMIT>
static void build_task(const Handle &h, bool upper, int q0, int q1, TaskScratch &S, TaskSizes &sz, uint2 *desc, unsigned char *stream,
                       size_t stream_off, unsigned *fdst, int *fsrc) {
  const int n = h.n, R = q1 - q0;
  const int *rows = h.lrows().data(), *cols = h.lcols().data(), *diag = h.ldiag().data();
  S.lev.resize(R); S.order.resize(R); S.slot.resize(R);
  int nl = 0;
  for (int q = q0; q < q1; ++q) {
    const int i = upper ? n - 1 - q : q;
    const int ps = upper ? diag[i] + 1 : rows[i], pe = upper ? rows[i + 1] : diag[i];
    int l = 0;
    for (int p = ps; p < pe; ++p) {
      const int qq = upper ? n - 1 - cols[p] : cols[p];
      if (qq >= q0) l = std::max(l, S.lev[qq - q0] + 1);
    }
    S.lev[q - q0] = l; nl = std::max(nl, l + 1);
  }
  S.cnt.assign(nl + 1, 0); S.stepb.assign(nl + 1, 0);
  for (int r = 0; r < R; ++r) S.cnt[S.lev[r] + 1]++;
  for (int l = 0; l < nl; ++l) S.stepb[l + 1] = S.stepb[l] + (S.cnt[l + 1] + 31) / 32;
  for (int l = 0; l < nl; ++l) S.cnt[l + 1] += S.cnt[l];          // cnt[l] = first position of level l in `order`
  {
    std::vector<int> &fill = S.slot;                               // reused below; here: running fill per level
    std::vector<int> pos(S.cnt.begin(), S.cnt.end() - 1);
    for (int r = 0; r < R; ++r) { const int l = S.lev[r]; S.order[pos[l]++] = r; }
    for (int l = 0; l < nl; ++l)
      for (int t = S.cnt[l]; t < S.cnt[l + 1]; ++t) { const int idx = t - S.cnt[l]; fill[S.order[t]] = (S.stepb[l] + idx / 32) * 32 + (idx & 31); }
  }
  const int nsteps = S.stepb[nl];
  sz.steps = nsteps;
  size_t cur = 0; long long nf = 0; int k = 0;
  for (int l = 0; l < nl; ++l) {
    for (int t0 = S.cnt[l]; t0 < S.cnt[l + 1]; t0 += 32, ++k) {
      const int nr = std::min(32, S.cnt[l + 1] - t0);
      int W = 0;
      for (int t = 0; t < nr; ++t) {
        const int i = upper ? n - 1 - (q0 + S.order[t0 + t]) : q0 + S.order[t0 + t];
        W = std::max(W, upper ? rows[i + 1] - diag[i] - 1 : diag[i] - rows[i]);
      }
      const int E = ru4(W * nr), nr4 = ru4(nr);
      const int bytes = 16 + 12 * E + (upper ? 12 : 4) * nr4;
      sz.maxblock = std::max(sz.maxblock, bytes); sz.maxw = std::max(sz.maxw, W);
      if (EMIT) {
        unsigned char *blk = stream + stream_off + cur;
        desc[k] = make_uint2((unsigned)((stream_off + cur) >> 4), (unsigned)(bytes >> 4));
        int *hd = (int *)blk; hd[0] = nr; hd[1] = W; hd[2] = E; hd[3] = 0;
        double *vals = (double *)(blk + 16);
        unsigned *codes = (unsigned *)(blk + 16 + 8 * (size_t)E);
        double *rowv = (double *)(blk + 16 + 12 * (size_t)E);
        int *rowid = (int *)(blk + 16 + 12 * (size_t)E + (upper ? 8 * (size_t)nr4 : 0));
        for (int e = 0; e < E; ++e) { vals[e] = 0.0; codes[e] = TT_PAD; }
        for (int t = 0; t < nr4; ++t) { rowid[t] = -1; if (upper) rowv[t] = 0.0; }
        for (int t = 0; t < nr; ++t) {
          const int i = upper ? n - 1 - (q0 + S.order[t0 + t]) : q0 + S.order[t0 + t];
          const int ps = upper ? diag[i] + 1 : rows[i], pe = upper ? rows[i + 1] : diag[i];
          rowid[t] = i;
          if (upper) { fdst[nf] = (unsigned)(((unsigned char *)(rowv + t) - stream) >> 3); fsrc[nf] = diag[i]; ++nf; }
          for (int p = ps; p < pe; ++p) {
            const int j = p - ps, c = cols[p];
            const int qq = upper ? n - 1 - c : c;
            unsigned code = (unsigned)c;
            if (qq >= q0) {
              const int ds = S.slot[qq - q0], age = k - (ds >> 5);
              if (age < TT_D) code = TT_LOCAL | ((unsigned)age << 8) | (unsigned)(ds & 31);
            }
            codes[j * nr + t] = code;
            fdst[nf] = (unsigned)(((unsigned char *)(vals + j * nr + t) - stream) >> 3); fsrc[nf] = p; ++nf;
          }
        }
      } else {
        for (int t = 0; t < nr; ++t) {
          const int i = upper ? n - 1 - (q0 + S.order[t0 + t]) : q0 + S.order[t0 + t];
          nf += (upper ? rows[i + 1] - diag[i] - 1 : diag[i] - rows[i]) + (upper ? 1 : 0);
        }
      }
      cur += bytes;
    }
  }
  sz.bytes = (long long)cur; sz.fills = nf;
}

// Task boundaries (sweep positions).  Rows are grouped the way a natural FE numbering is built:
//   line  = maximal run of rows in which every row needs its predecessor (a lane follows one line);
//   plane = maximal run of lines in which every line needs the line before it.
// A task is a group of up to 32 consecutive lines of ONE plane, and every plane is cut the same way.
// The alignment matters: a row's operands in the previous plane are then produced at (nearly) the
// same local step in their own task, so consecutive planes pipeline with a lag of two L2 round trips;
// with boundaries drifting from plane to plane the lag grows to tens of steps per plane (measured:
// 328939 steps, makespan 72397 steps on the 201^3 grid with uniform ranges, 1338 with aligned ones).
static std::vector<int> task_bounds(const Handle &h, bool upper, int fixed_rows) {
  const int n = h.n;
  std::vector<int> bounds;
  bounds.push_back(0);
  if (fixed_rows > 0) { for (long long q = fixed_rows; q < n; q += fixed_rows) bounds.push_back((int)q); bounds.push_back(n); return bounds; }
  const int MAXLINE = 4096, MINROWS = 512;
  auto maxdep = [&](int q) -> int {                       // highest sweep position row q reads (-1: none)
    const int i = upper ? n - 1 - q : q, d = h.ldiag()[i];
    if (!upper) return d > h.lrows()[i] ? h.lcols()[d - 1] : -1;
    return d + 1 < h.lrows()[i + 1] ? n - 1 - h.lcols()[d + 1] : -1;
  };
  std::vector<int> ls;                                    // line starts; bit 30 marks a plane start
  std::vector<char> ps;
  int prev_start = 0, len = 0;
  for (int q = 0; q < n; ++q) {
    const int md = maxdep(q);
    const bool start = q == 0 || md != q - 1 || len >= MAXLINE;
    if (start) { ls.push_back(q); ps.push_back(q == 0 || md < prev_start); prev_start = q; len = 0; }
    ++len;
  }
  const int nlines = (int)ls.size();
  ls.push_back(n);
  int G = 32, in_task = 0, task_rows = 0;
  for (int k = 0; k < nlines; ++k) {
    if (ps[k]) {
      int e = k + 1; while (e < nlines && !ps[e]) ++e;
      const int nl = e - k, ng = (nl + 31) / 32;
      if (task_rows >= MINROWS) { bounds.push_back(ls[k]); in_task = 0; task_rows = 0; }
      G = (nl + ng - 1) / ng;
    }
    if (in_task >= G) { bounds.push_back(ls[k]); in_task = 0; task_rows = 0; }
    ++in_task; task_rows += ls[k + 1] - ls[k];
  }
  bounds.push_back(n);
  return bounds;
}

static void build_plan(Handle &h, TriTask &P, bool upper, int fixed_rows) {
  const std::vector<int> tb = task_bounds(h, upper, fixed_rows);
  P.ntasks = (int)tb.size() - 1; P.upper = upper; P.rows_per_task = P.ntasks ? h.n / P.ntasks : 0;
  std::vector<TaskSizes> sz(P.ntasks);
  int nth = omp_get_max_threads();
  std::vector<TaskScratch> scratch(nth);
#pragma omp parallel for schedule(dynamic, 4)
  for (int t = 0; t < P.ntasks; ++t)
    build_task<false>(h, upper, tb[t], tb[t + 1], scratch[omp_get_thread_num()], sz[t], nullptr, nullptr, 0, nullptr, nullptr);
  std::vector<int> step0(P.ntasks + 1, 0);
  std::vector<size_t> boff(P.ntasks + 1, 0); std::vector<long long> foff(P.ntasks + 1, 0);
  P.slot_bytes = 16; P.maxw = 0;
  for (int t = 0; t < P.ntasks; ++t) {
    B200_REQUIRE((long long)step0[t] + sz[t].steps < 2147483647LL, "task plan: too many steps");
    step0[t + 1] = step0[t] + (int)sz[t].steps; boff[t + 1] = boff[t] + (size_t)sz[t].bytes; foff[t + 1] = foff[t] + sz[t].fills;
    P.slot_bytes = std::max(P.slot_bytes, sz[t].maxblock); P.maxw = std::max(P.maxw, sz[t].maxw);
  }
  P.nsteps = step0[P.ntasks]; P.nbytes = boff[P.ntasks]; P.nfill = foff[P.ntasks];
  B200_REQUIRE((P.nbytes >> 4) < 4294967295ULL, "task plan: stream exceeds 64 GB");
  std::vector<uint2> desc(std::max<long long>(P.nsteps, 1));
  unsigned char *stream = (unsigned char *)malloc(std::max<size_t>(P.nbytes, 16));
  std::vector<unsigned> fdst(std::max<long long>(P.nfill, 1)); std::vector<int> fsrc(std::max<long long>(P.nfill, 1));
  B200_REQUIRE(stream != nullptr, "task plan: host allocation failed");
#pragma omp parallel for schedule(dynamic, 4)
  for (int t = 0; t < P.ntasks; ++t) {
    TaskSizes s2;
    build_task<true>(h, upper, tb[t], tb[t + 1], scratch[omp_get_thread_num()], s2, desc.data() + step0[t], stream, boff[t],
                     fdst.data() + foff[t], fsrc.data() + foff[t]);
  }
  cudaStream_t st = h.stream;
  P.step0.ensure(P.ntasks + 1); P.desc.ensure(desc.size()); P.stream.ensure(std::max<size_t>(P.nbytes, 16));
  P.fill_dst.ensure(fdst.size()); P.fill_src.ensure(fsrc.size());
  B200_CUDA(cudaMemcpyAsync(P.step0.p, step0.data(), step0.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  B200_CUDA(cudaMemcpyAsync(P.desc.p, desc.data(), desc.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
  if (P.nbytes) B200_CUDA(cudaMemcpyAsync(P.stream.p, stream, P.nbytes, cudaMemcpyHostToDevice, st));
  if (P.nfill) {
    B200_CUDA(cudaMemcpyAsync(P.fill_dst.p, fdst.data(), (size_t)P.nfill * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(P.fill_src.p, fsrc.data(), (size_t)P.nfill * sizeof(int), cudaMemcpyHostToDevice, st));
  }
  B200_CUDA(cudaStreamSynchronize(st));
  free(stream);
}

void tritask_release(Handle &h) {
  for (TriTask *P : {&h.TL, &h.TU}) { P->step0.release(); P->desc.release(); P->stream.release(); P->fill_dst.release(); P->fill_src.release(); P->ntasks = 0; }
  h.d_ytask.release(); h.d_xtask.release(); h.tt_ready = false;
}

void tritask_analyse(Handle &h) {
  if (h.tt_ready) return;
  build_plan(h, h.TL, false, h.tt_rows);
  build_plan(h, h.TU, true, h.tt_rows);
  h.d_ytask.ensure(std::max(h.n, 1)); h.d_xtask.ensure(std::max(h.n, 1));
  h.tt_ready = true;
  if (getenv("B200_TT_DEBUG"))
    for (const TriTask *P : {&h.TL, &h.TU})
      fprintf(stderr, "[tritask] %s rows/task %d tasks %d steps %lld (%.1f rows/step) stream %.3f GB slot %d B maxw %d\n", P->upper ? "U" : "L",
              P->rows_per_task, P->ntasks, P->nsteps, P->nsteps ? (double)h.n / P->nsteps : 0.0, P->nbytes / 1e9, P->slot_bytes, P->maxw);
}

// -----------------------------------------------------------------------------------------------
__global__ void k_tt_fill(long long nfill, const unsigned *__restrict__ dst, const int *__restrict__ src, const double *__restrict__ lu,
                          double *__restrict__ stream) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nfill; i += (long long)gridDim.x * blockDim.x)
    stream[dst[i]] = lu[src[i]];
}
void tritask_refresh_values(Handle &h) {
  for (TriTask *P : {&h.TL, &h.TU})
    if (P->nfill) k_tt_fill<<<NUM_SMS * 16, 256, 0, h.stream>>>(P->nfill, P->fill_dst.p, P->fill_src.p, h.d_ilu.p, (double *)P->stream.p);
  B200_CUDA(cudaGetLastError());
}

// -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem, const void *g) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// guarded load of a result-vector entry: predicated in PTX so that no branch is generated
__device__ __forceinline__ double ld_relaxed_if(const double *p, bool pred) {
  double v = 0.0;
  asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q ld.relaxed.gpu.global.f64 %0, [%1]; }" : "+d"(v) : "l"(p), "r"((unsigned)pred));
  return v;
}

struct TaskView {
  int ntasks; const int *step0; const uint2 *desc; const unsigned char *stream; int slot_bytes;
};
struct StepRegs { double x[TT_CH]; double rhs; int row; };

// forward (UPPER = false): out_i = rhs_i - sum_{j<i} L_ij out_j            CRSMatrix.F90:4642-4649
// backward (UPPER = true): out_i = Dinv_i (rhs_i - sum_{j>i} U_ij out_j)   CRSMatrix.F90:4653-4660
template <bool UPPER, int P1>
__global__ void __launch_bounds__(256, 1) k_tritask(TaskView T, const double *__restrict__ rhs, double *out, double *out2, Ctrl *ctrl, unsigned wait_ns, int nrows, int pf_dist, long long *trace) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (ctrl->done) return;
  constexpr int NS = P1 + 1;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int NW = gridDim.x * wpb, vw = wib * gridDim.x + blockIdx.x;   // consecutive tasks land on different SMs
  const size_t per_warp = (size_t)TT_D * 256 + (size_t)NS * T.slot_bytes;
  double *ring = (double *)(smem_raw + wib * per_warp);
  unsigned char *stage = smem_raw + wib * per_warp + TT_D * 256;
  long long spins = 0;

  for (int task = vw; task < T.ntasks; task += NW) {
    const int s0 = T.step0[task], nst = T.step0[task + 1] - s0;
    long long t_start = 0, n_rounds = 0, n_sleeps = 0;
    if (trace) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    int kb = 0;
    uint2 dcur = make_uint2(0, 0), dnxt = make_uint2(0, 0);
    if (lane < nst) dcur = __ldg(T.desc + s0 + lane);
    if (32 + lane < nst) dnxt = __ldg(T.desc + s0 + 32 + lane);
    int slot_issue = 0;
    auto issue = [&](int k) {                                  // stage block k (uniform k)
      if (k < nst) {
        const int r = k - kb;
        const unsigned ax = __shfl_sync(0xffffffffu, dcur.x, r & 31), ay = __shfl_sync(0xffffffffu, dcur.y, r & 31);
        const unsigned bx = __shfl_sync(0xffffffffu, dnxt.x, r & 31), by = __shfl_sync(0xffffffffu, dnxt.y, r & 31);
        const unsigned off16 = r < 32 ? ax : bx, n16 = r < 32 ? ay : by;
        const unsigned char *src = T.stream + ((size_t)off16 << 4);
        unsigned char *dst = stage + (size_t)slot_issue * T.slot_bytes;
        for (unsigned p = lane; p < n16; p += 32) cp_async16(dst + (p << 4), src + ((size_t)p << 4));
      }
      cp_async_commit();
      slot_issue = slot_issue + 1 == NS ? 0 : slot_issue + 1;
    };
    // Both stages are written branch-free: lanes without a row mirror lane nr-1 (every address stays inside
    // the block), guarded global loads are predicated in PTX, and the choice ring / result vector is a
    // select.  (A first version with `if` around every load spent ~4000 cycles per step in divergence
    // reconvergence: 141 BSSY/BSYNC pairs in the loop body.)
    auto gather = [&](int slot, StepRegs &R) {                 // request rhs and the remote operands of a staged step
      const unsigned char *blk = stage + (size_t)slot * T.slot_bytes;
      const int4 hd = *(const int4 *)blk;
      const int nr = hd.x, W = hd.y, E = hd.z, nr4 = (nr + 3) & ~3;
      const unsigned *codes = (const unsigned *)(blk + 16 + 8 * (size_t)E);
      const int *rowid = (const int *)(blk + 16 + 12 * (size_t)E + (UPPER ? 8 * (size_t)nr4 : 0));
      const bool act = lane < nr;
      const int ln = min(lane, nr - 1);
      R.row = rowid[ln];
      R.rhs = __ldg(rhs + R.row);
      {                                                        // the lane follows a run of consecutive rows: pull the run's rhs into L2 ahead
        const double *pf = rhs + (UPPER ? max(R.row - pf_dist, 0) : min(R.row + pf_dist, nrows - 1));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
      }
      const int wl = max(W - 1, 0);
      unsigned c[TT_CH];
#pragma unroll
      for (int j = 0; j < TT_CH; ++j) c[j] = codes[min(j, wl) * nr + ln];
#pragma unroll
      for (int j = 0; j < TT_CH; ++j) R.x[j] = ld_relaxed_if(out + c[j], act && j < W && !(c[j] & TT_LOCAL));
    };
    auto compute = [&](int k, int slot, const StepRegs &R) {
      const unsigned char *blk = stage + (size_t)slot * T.slot_bytes;
      const int4 hd = *(const int4 *)blk;
      const int nr = hd.x, W = hd.y, E = hd.z;
      const double *vals = (const double *)(blk + 16);
      const unsigned *codes = (const unsigned *)(blk + 16 + 8 * (size_t)E);
      const double *rowv = (const double *)(blk + 16 + 12 * (size_t)E);
      const bool act = lane < nr;
      const int ln = min(lane, nr - 1);
      double s = R.rhs;
      for (int c0 = 0; c0 < W; c0 += TT_CH) {
        // (the asm loads carry no memory clobber and sit in their own loop: a clobbering asm between the
        //  shared-memory loads serialised the 16 entries, ~160 cycles each)
        unsigned cd[TT_CH]; double v[TT_CH], x[TT_CH], xr[TT_CH];
        unsigned use = 0, pend = 0;
#pragma unroll
        for (int j = 0; j < TT_CH; ++j) {
          const int jc = min(c0 + j, W - 1);
          cd[j] = codes[jc * nr + ln];
          v[j] = vals[jc * nr + ln];
        }
#pragma unroll
        for (int j = 0; j < TT_CH; ++j) {
          x[j] = ring[((k - (int)((cd[j] >> 8) & 0xFFu)) & (TT_D - 1)) * 32 + (cd[j] & 31u)];
          use |= (unsigned)(act && c0 + j < W && cd[j] != TT_PAD) << j;
        }
        if (c0 == 0) {
#pragma unroll
          for (int j = 0; j < TT_CH; ++j) xr[j] = R.x[j];
        } else {
#pragma unroll
          for (int j = 0; j < TT_CH; ++j) xr[j] = ld_relaxed_if(out + cd[j], ((use >> j) & 1u) && !(cd[j] & TT_LOCAL));
        }
#pragma unroll
        for (int j = 0; j < TT_CH; ++j) {
          const bool loc = (cd[j] & TT_LOCAL) != 0;
          x[j] = loc ? x[j] : xr[j];
          pend |= (unsigned)(((use >> j) & 1u) && !loc && is_sentinel(x[j])) << j;
        }
        // Operands still unset: re-read them.  After two failed rounds the warp is ahead of its producers
        // (typically a task whose turn has not come): one lane then sleeps on ONE address, so that the
        // hundreds of waiting warps do not flood L2 with polls and slow the warps on the critical path.
        for (int round = 0; __any_sync(0xffffffffu, pend != 0); ++round) {
          ++n_rounds;
          if (round >= 2) {
            ++n_sleeps;
            const unsigned bal = __ballot_sync(0xffffffffu, pend != 0);
            if (lane == __ffs(bal) - 1) {
              const int j0 = __ffs(pend) - 1;
              unsigned c = 0;
#pragma unroll
              for (int j = 0; j < TT_CH; ++j) if (j == j0) c = cd[j];
              while (is_sentinel(ld_relaxed(out + c))) {
                if (++spins > TT_SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
                __nanosleep(wait_ns);
              }
            }
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < TT_CH; ++j)
            if (pend & (1u << j)) { x[j] = ld_relaxed(out + cd[j]); if (!is_sentinel(x[j])) pend &= ~(1u << j); }
          if (++spins > TT_SPIN_LIMIT) { ctrl->spin_timeout = 1; pend = 0; }
        }
#pragma unroll
        for (int j = 0; j < TT_CH; ++j) { const double t = nfms(s, v[j], x[j]); s = (use >> j) & 1u ? t : s; }
      }
      if (act) {
        double res = UPPER ? __dmul_rn(rowv[lane], s) : s;
        if (res != res) res = __longlong_as_double((long long)CANON_NAN);
        ring[(k & (TT_D - 1)) * 32 + lane] = res;
        st_relaxed(out + R.row, res);
        if (out2) out2[R.row] = res;
      }
    };

    for (int k = 0; k < P1; ++k) issue(k);
    cp_async_wait<P1 - 1>();
    __syncwarp();
    StepRegs RC, RN;
    int slot_cur = 0;
    gather(0, RC);
    RN = RC;
    for (int k = 0; k < nst; ++k) {
      if (k - kb >= 32) { kb += 32; dcur = dnxt; dnxt = make_uint2(0, 0); if (kb + 32 + lane < nst) dnxt = __ldg(T.desc + s0 + kb + 32 + lane); }
      issue(k + P1);
      cp_async_wait<P1 - 1>();
      __syncwarp();
      const int slot_nxt = slot_cur + 1 == NS ? 0 : slot_cur + 1;
      if (k + 1 < nst) gather(slot_nxt, RN);
      compute(k, slot_cur, RC);
      __syncwarp();
      slot_cur = slot_nxt;
      RC = RN;
    }
    cp_async_wait<0>();
    __syncwarp();
    if (trace && lane == 0) {
      long long t_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
      trace[4 * task + 0] = t_start; trace[4 * task + 1] = t_end; trace[4 * task + 2] = n_rounds; trace[4 * task + 3] = n_sleeps;
    }
  }
}

__global__ void k_tt_prepare(int n, double *a, double *b) {
  const double sent = __longlong_as_double((long long)SENTINEL);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { a[i] = sent; b[i] = sent; }
}

template <bool UPPER, int P1>
static void tt_launch(Handle &h, const TriTask &P, int wpb, const double *rhs, double *out, double *out2) {
  const size_t smem = (size_t)wpb * ((size_t)TT_D * 256 + (size_t)(P1 + 1) * P.slot_bytes);
  const void *kern = (const void *)k_tritask<UPPER, P1>;
  B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TaskView T; T.ntasks = P.ntasks; T.step0 = P.step0.p; T.desc = P.desc.p; T.stream = P.stream.p; T.slot_bytes = P.slot_bytes;
  Ctrl *ctrl = h.ctrl.p;
  int dev = 0, sms = 0;
  B200_CUDA(cudaGetDevice(&dev));
  B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int blocks = std::max(1, std::min(sms, (P.ntasks + wpb - 1) / wpb));
  unsigned wait_ns = h.tt_wait_ns; int nrows = h.n, pf_dist = h.tt_pf;
  long long *trace = nullptr;
  static DBuf<long long> tbuf;
  const char *tf = getenv("B200_TT_TRACE");      // debugging aid: per-task start/end stamps of the forward sweep
  if (tf && *tf && !UPPER) { tbuf.ensure((size_t)P.ntasks * 4); trace = tbuf.p; }
  void *argv[] = {(void *)&T, (void *)&rhs, (void *)&out, (void *)&out2, (void *)&ctrl, (void *)&wait_ns, (void *)&nrows, (void *)&pf_dist, (void *)&trace};
  B200_CUDA(cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(wpb * 32), argv, smem, h.stream));
  if (trace) {
    std::vector<long long> ht((size_t)P.ntasks * 4);
    B200_CUDA(cudaStreamSynchronize(h.stream));
    B200_CUDA(cudaMemcpy(ht.data(), trace, ht.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<int> st(P.ntasks + 1);
    B200_CUDA(cudaMemcpy(st.data(), P.step0.p, st.size() * sizeof(int), cudaMemcpyDeviceToHost));
    FILE *f = fopen(tf, "w");
    B200_REQUIRE(f != nullptr, "cannot open B200_TT_TRACE file");
    long long t0 = ht[0];
    for (int t = 0; t < P.ntasks; ++t) t0 = std::min(t0, ht[4 * t]);
    for (int t = 0; t < P.ntasks; ++t)
      fprintf(f, "%d steps %d start_us %.2f end_us %.2f poll_rounds %lld sleeps %lld\n", t, st[t + 1] - st[t], (ht[4 * t] - t0) / 1e3, (ht[4 * t + 1] - t0) / 1e3, ht[4 * t + 2], ht[4 * t + 3]);
    fclose(f);
  }
}

// largest warp count per block whose staging fits in shared memory (0: task mode not usable)
template <int P1> static int tt_fit(const Handle &h, int want) {
  const size_t budget = 220 * 1024;
  const size_t per_warp = (size_t)TT_D * 256 + (size_t)(P1 + 1) * std::max(h.TL.slot_bytes, h.TU.slot_bytes);
  int w = (int)std::min<size_t>(want, budget / per_warp);
  return std::min(w, 8);
}

bool tritask_usable(Handle &h) {
  if (!h.tt_ready || h.n == 0) return false;
  return tt_fit<2>(h, 1) >= 1;
}

void lu_apply_task(Handle &h, double *u, const double *v) {
  B200_REQUIRE(h.tt_ready, "task-mode triangular solve without a plan");
  double *xo = (u == v) ? h.d_xtask.p : u;
  double *x2 = (u == v) ? u : nullptr;
  k_tt_prepare<<<std::min((h.n + 255) / 256, NUM_SMS * 8), 256, 0, h.stream>>>(h.n, h.d_ytask.p, xo);
  const int want = h.tt_wpb > 0 ? h.tt_wpb : 4;
  if (tt_fit<4>(h, want) >= std::min(want, 2)) {
    const int wpb = tt_fit<4>(h, want);
    tt_launch<false, 4>(h, h.TL, wpb, v, h.d_ytask.p, nullptr);
    tt_launch<true, 4>(h, h.TU, wpb, h.d_ytask.p, xo, x2);
  } else {
    const int wpb = std::max(1, tt_fit<2>(h, want));
    tt_launch<false, 2>(h, h.TL, wpb, v, h.d_ytask.p, nullptr);
    tt_launch<true, 2>(h, h.TU, wpb, h.d_ytask.p, xo, x2);
  }
  B200_CUDA(cudaGetLastError());
  h.st_launch += 3; h.st_pcond++;
}

}  // namespace b200

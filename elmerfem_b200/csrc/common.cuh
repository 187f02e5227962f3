// Shared declarations of the B200 linear-solve library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>
#include "skewgeom.h"
#include "wavegeom.h"
#include "lanegeom.h"

namespace b200 {

// ------------------------------------------------------------------------------------------
// errors: every failure throws; the C ABI layer turns it into a return code + message.
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
void set_last_error(const std::string &s);

#define B200_CUDA(x)                                                                            \
  do {                                                                                          \
    cudaError_t e__ = (x);                                                                      \
    if (e__ != cudaSuccess) {                                                                   \
      char m__[512];                                                                            \
      snprintf(m__, sizeof m__, "CUDA error '%s' in %s at %s:%d", cudaGetErrorString(e__), #x,   \
               __FILE__, __LINE__);                                                             \
      throw b200::Error(m__);                                                                   \
    }                                                                                           \
  } while (0)
#define B200_REQUIRE(cond, msg)                                                                 \
  do { if (!(cond)) throw b200::Error(std::string("elmer_b200: ") + (msg)); } while (0)

// ------------------------------------------------------------------------------------------
// constants shared with the reference
constexpr double AEPS = 10.0 * 2.220446049250313e-16;      // Types.F90:71
constexpr double HUTI_EPSILON = 1.17549435E-38;            // huti_fdefs.h:14
constexpr int SLICE = 32;                                  // SELL slice height = warp width
constexpr int NUM_SMS = 148;                               // B200
// "not yet computed" marker of the sync-free triangular solves (a NaN payload no arithmetic produces)
constexpr unsigned long long SENTINEL = 0x7FF4DEADBEEF0B20ULL;
constexpr unsigned long long CANON_NAN = 0x7FF8000000000000ULL;

// ------------------------------------------------------------------------------------------
// device buffers
template <class T> struct DBuf {
  T *p = nullptr; size_t cap = 0;
  void ensure(size_t n) {
    if (n <= cap) return;
    release();
    B200_CUDA(cudaMalloc((void **)&p, (n ? n : 1) * sizeof(T)));
    cap = n ? n : 1;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Sliced-ELL view: slot = slice*32 + lane.  Entry j of a slot sits at ptr[slice] + j*32 + lane.
// perm == nullptr: slot == row (SpMV operand).  Otherwise slot -> row (-1 = empty), level order.
struct SellView {
  int nslots = 0, nslices = 0;
  const long long *ptr = nullptr;   // nslices+1, in entries
  const int *len = nullptr;         // per slot
  const int *perm = nullptr;        // per slot or null
  const int *cols = nullptr;        // 0-based local column ids
  const double *vals = nullptr;
};

struct Sell {
  int nslots = 0, nslices = 0; long long nstore = 0;
  DBuf<long long> ptr; DBuf<int> len; DBuf<int> perm; DBuf<int> cols; DBuf<double> vals;
  DBuf<int> start;                  // per slot: position in the CRS value array of entry 0
  DBuf<int> gate;                   // per slice (level plans): the slice's level
  bool has_perm = false, ralign = false;   // ralign: rows right-aligned inside their slice (see k_sell_fill)
  SellView view() const {
    SellView v; v.nslots = nslots; v.nslices = nslices; v.ptr = ptr.p; v.len = len.p;
    v.perm = has_perm ? perm.p : nullptr; v.cols = cols.p; v.vals = vals.p; return v;
  }
  void release() { ptr.release(); len.release(); perm.release(); cols.release(); vals.release(); start.release(); gate.release(); }
};

// device-resident control block of a running Krylov solve
struct Ctrl {
  int done;        // 1 = stop: every kernel of the iteration returns immediately
  int info;        // HUTI_INFO
  int iters;       // HUTI_ITERS
  int flag;        // method specific (BiCGStab: |s| < HUTI_EPSILON early exit pending)
  int spin_timeout;// a sync-free kernel gave up waiting (reported as HUTI_HALTED)
  int pad[3];
  double residual;
  double tol, maxtol, bnorm;
  int maxit, minit, stopc, pad2;
};

// reduction scratch: NRED independent reductions, each with per-block partials
constexpr int MAX_RED_BLOCKS = NUM_SMS * 16;
constexpr int NRED = 24;            // enough for the (l+1)(l+2)/2 = 21 Gram dots of BiCGStab(l=5)
constexpr int NSCAL = 64;

struct Halo;   // multi-GPU plan (comm.cu)

// wave-tile triangular solve (wave.cu, B200_TRI_MODE=3): geometry, tile tables, the two per-step entry streams, work vectors
struct WavePlan {
  WaveGeom g; DBuf<int> tile_of, tile_sig, tile_grp; DBuf<double> SL, SU, yin, y, x; DBuf<long long> trace;
  DBuf<int> mapL, mapU;            // per stream entry: position of its value in the ILU array, -1 = pad (built once per structure)
  bool ready = false, tried = false, trace_on = false;
};

// lane-tile triangular solve (lane.cu, B200_TRI_MODE=4): geometry, tile tables, the two per-step entry streams, work vectors
struct LanePlan {
  LaneGeom g; DBuf<int> tile_of, tile_sig, tile_grp; DBuf<double> SL, SU, y, x; DBuf<long long> trace;
  DBuf<int> mapL, mapU;            // as in WavePlan
  bool ready = false, tried = false, trace_on = false, traced = false;
};

struct Handle {
  int device = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev_end = nullptr, evf0 = nullptr, evf1 = nullptr;
  // structure
  int n = 0; long long nnz = 0; int ndeg = 1; int index_base = 1; int nacc = 1; bool nacc_blocked = false;
  std::vector<int> h_rows, h_cols, h_diag;       // 0-based host copies (level analysis)
  DBuf<int> d_rows_in, d_cols_in, d_diag_in;     // bit-exact mirrors of the caller's arrays
  DBuf<int> d_rows, d_cols, d_diag;              // 0-based working copies
  // values
  DBuf<double> d_vals, d_prec, d_ilu, d_dvals; bool have_vals = false, have_prec = false, ilu_valid = false, ilu_exists = false;
  // ILU(n > 0): the factor lives on its own pattern (CRSMatrix.F90:3488-3510); ILU0 aliases the matrix pattern
  int ilu_order = 0; bool ilu_pat_ready = false; long long ilu_nnz = 0;
  int bilu_blocks = 0;                             // > 1: BILU, the factor of the block-diagonal part (CRS_BlockDiagonal, CRSMatrix.F90:2382-2420)
  bool ilut = false; double ilut_tol = 0.0;      // ILUT: the pattern comes out of the factorisation (ilut_factor installs it)
  bool ilu_sep() const { return ilu_order > 0 || bilu_blocks > 1 || ilut; }   // factor on its own pattern
  std::vector<int> hl_rows, hl_cols, hl_diag;    // 0-based host copies of ILURows/ILUCols/ILUDiag
  DBuf<int> dl_rows, dl_cols, dl_diag, dl_src;   // device copies; dl_src: position of the entry in the matrix values, -1 for fill
  const std::vector<int> &lrows() const { return ilu_sep() ? hl_rows : h_rows; }
  const std::vector<int> &lcols() const { return ilu_sep() ? hl_cols : h_cols; }
  const std::vector<int> &ldiag() const { return ilu_sep() ? hl_diag : h_diag; }
  const int *d_lrows() const { return ilu_sep() ? dl_rows.p : d_rows.p; }
  const int *d_lcols() const { return ilu_sep() ? dl_cols.p : d_cols.p; }
  const int *d_ldiag() const { return ilu_sep() ? dl_diag.p : d_diag.p; }
  long long lnnz() const { return ilu_sep() ? ilu_nnz : nnz; }
  // device-side Linear System Scaling (b200_scale_system): D, and D * bnorm of the running solve
  bool scaled = false; DBuf<double> d_scale, d_scale_rhs;
  // SpMV operand
  Sell A;
  // ILU0 + triangular solves
  bool tri_ready = false; int nlev_f = 0, nlev_b = 0;
  DBuf<int> d_order_f; DBuf<int> d_rowdone;      // factorisation order (rows sorted by forward level)
  std::vector<int> h_level_f;
  Sell L, U; DBuf<double> d_dinv_slot;           // U slots carry the inverted diagonal
  int tri_node = 0; bool tri_node_u = false, tri_node_off = false;    // tri_node > 0: L/U plans are in the node-lane layout with this many dofs per node (structure.cu); off: never use it
  DBuf<int> tri_counters; int tri_maxw = 0, tri_lookahead = 2; unsigned tri_gate_sleep = 100, tri_spin_sleep = 0;
  DBuf<int> d_lvlcnt_f, d_lvlcnt_b;              // slices per level (forward / backward)
  DBuf<int> d_urhs; DBuf<double> d_yl, d_xu;     // backward rhs map (U slot -> L slot); slot-ordered solve vectors
  // incomplete Cholesky ('Linear System Symmetric ILU', A % Cholesky): the factor's lower part by columns (rows descending: the order
  // in which CRS_LUSolve's column-oriented backward loop updates an unknown) and the level plan of that sweep
  int ilu_reg_ok = -1;                           // rows narrow enough for k_ilu0_factor_reg: -1 unknown, 0 no, 1 yes
  bool cholesky = false, ch_ready = false; int ch_nlev = 0, ch_nslices = 0;
  DBuf<int> ch_ptr, ch_row, ch_pos, ch_perm, ch_gate, ch_lvlcnt, ch_counters; DBuf<double> ch_y, ch_x;
  // tri_mode: 0 level kernel, 3 wave tiles, 4 lane tiles, -2 time them at the first factorisation and keep the fastest
  WavePlan wv; int wv_blocks_per_sm = 0, wv_cfg = 0, wv_e = 3;
  LanePlan lt; int lt_tc = 1, lt_warps = 0, lt_e = 1;
  void *stage_buf[2] = {nullptr, nullptr}; cudaEvent_t stage_ev[2] = {nullptr, nullptr}; bool stage_uploads = true;   // pinned bounce buffers of b200_set_values
  bool mv_honor_skip = false;                     // partitioned SpMVs queued inside a conditional section (Ctrl::done == 2) return at once
  bool bl_host = false;                           // B200_BICGSTABL_HOST=1: host-driven BiCGStab(l) (the round-1 driver) instead of the device-resident one
  int tri_mode = 0, tri_mode_cfg = 0;
  // workspace
  std::vector<DBuf<double>> work; DBuf<double> d_b, d_x, d_tmp, d_P;
  DBuf<double> red_partials; DBuf<unsigned int> red_counters; DBuf<double> scal; DBuf<Ctrl> ctrl;
  double *h_pinned = nullptr; Ctrl *h_ctrl = nullptr;   // pinned host mirrors (h_pinned: NSCAL doubles, h_ctrl: 2 slots)
  // multi-GPU
  Halo *halo = nullptr; void *nccl = nullptr; int nranks = 1, rank = 0; long long gn = 0;
  // stats
  double st_solve_ms = 0, st_factor_ms = 0, st_spmv_ms = 0, st_lu_ms = 0, st_resid = 0;
  long long st_matvec = 0, st_pcond = 0, st_launch = 0, st_launch_last = 0, st_h2d = 0, st_d2h = 0, st_iters = 0, st_factor_launch = 0;
  // tuning (env overridable)
  int spmv_blocks = 0, tri_blocks_per_sm = 0, blas_blocks = NUM_SMS * 8;
  int grid_ilu = 0, grid_tri_l = 0, grid_tri_u = 0;   // co-resident grid sizes (occupancy x SMs)
  const void *grid_ilu_kern = nullptr;
  DBuf<unsigned char> d_ilu_pos; DBuf<long long> d_ilu_posptr; int ilu_map_maxu = 0; bool ilu_map_tried = false;   // position map of the elimination (precond.cu)
  // B200_PIN_VALUES=1: the caller's value array is page-locked (cudaHostRegister) the first time it is seen, so the
  // once-per-nonlinear-iteration upload runs at PCIe speed instead of through a staging copy
  int pin_values = 0; const void *pinned_ptr = nullptr; size_t pinned_bytes = 0;
  // hook-1 state
  const double *hook_vals_ptr = nullptr; double hook_checksum = 0;
};

// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ double ld_stream(const double *p) {
  double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}
__device__ __forceinline__ int ld_stream(const int *p) {
  int v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ double ld_relaxed(const double *p) {
  double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_relaxed(double *p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire(const int *p) {
  int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool is_sentinel(double v) {
  return (unsigned long long)__double_as_longlong(v) == SENTINEL;
}
// a - b*c and a + b*c with separate roundings: the reference is compiled without FMA contraction
__device__ __forceinline__ double nfms(double a, double b, double c) { return __dsub_rn(a, __dmul_rn(b, c)); }
__device__ __forceinline__ double nfma(double a, double b, double c) { return __dadd_rn(a, __dmul_rn(b, c)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic grid reduction of NV values per thread.  Each block writes one partial per value,
// the last block to arrive (ticket) sums the partials in a fixed order and hands the totals to
// `fin(totals)` executed by thread 0 of that block.  Block size must be a multiple of 32, <= 1024.
template <int NV, class Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], double *partials, unsigned int *counter, Fin fin) {
  __shared__ double sm[NV][32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) { double s = warp_sum(v[k]); if (lane == 0) sm[k][warp] = s; }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = lane < nw ? sm[k][lane] : 0.0;
      s = warp_sum(s);
      if (lane == 0) partials[(size_t)k * MAX_RED_BLOCKS + blockIdx.x] = s;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double tot[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(partials + (size_t)k * MAX_RED_BLOCKS + i);
    s = warp_sum(s);
    if (lane == 0) sm[k][warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = lane < nw ? sm[k][lane] : 0.0;
      tot[k] = warp_sum(s);
    }
    if (lane == 0) { *counter = 0u; fin(tot); }
  }
}
#endif

// ------------------------------------------------------------------------------------------
// host-side entry points of the kernel translation units
void structure_build(Handle &h);                       // SELL operand from the CRS mirror
void sell_refresh_values(Handle &h, Sell &S, const double *crs_vals);
void sell_finish(Handle &h, Sell &S, int nslots, bool has_perm, const int *src_cols);   // start/len/perm pre-filled
void ilu_pattern_build(Handle &h);                     // ILU(n > 0): symbolic fill (host), once per structure and order
void ilu_invalidate(Handle &h);                        // forget factor, plans and ILU(n) pattern (structure or order changed)
void tri_analyse(Handle &h);                           // levels + L/U level-sorted SELL plans
void ilu0_factor(Handle &h);                           // d_ilu from d_prec/d_vals, refresh L/U values
void lu_apply(Handle &h, double *u, const double *v);  // u = (LU)^-1 v   (device pointers)
void diag_apply(Handle &h, double *u, const double *v);
void sgs_sweeps(Handle &h, const double *b, double *x, double *t1, double *t2, double omega);   // one forward + one backward Gauss-Seidel sweep
// stream refill by gather (structure.cu): S[e] = map[e] >= 0 ? ilu[map[e]] : 0, and the one-time construction of the map from a scatter of indices
void stream_map_build(Handle &h, long long n, const double *S_with_indices, int *map);
void stream_gather(Handle &h, long long n, const int *map, const double *ilu, double *S);
void stream_iota1(Handle &h, long long n, double *a);
void wave_analyse(Handle &h);                          // wave-tile plan (host detection of the grid stencil; no-op when it does not apply)
void wave_refresh_values(Handle &h);
void wave_release(Handle &h);
void lu_apply_wave(Handle &h, double *u, const double *v);
void ichol_analyse(Handle &h);                         // column lists + backward level plan of the incomplete Cholesky solve
void ichol_release(Handle &h);
void lane_analyse(Handle &h);                          // lane-tile plan (same detection; no-op when it does not apply)
void lane_refresh_values(Handle &h);
void lane_release(Handle &h);
void lu_apply_lane(Handle &h, double *u, const double *v);
void lane_trace_enable(Handle &h, bool on);
void lane_trace_fetch(Handle &h, std::vector<long long> &out);
void wave_trace_enable(Handle &h, bool on);
void wave_trace_fetch(Handle &h, std::vector<long long> &out);
void halo_release(Handle &h);
size_t vec_len(const Handle &h);                       // n + ghost entries: length every SpMV operand must have
void matvec_full(Handle &h, const double *x, double *y);   // y = A x incl. halo exchange when partitioned
void install_structure(Handle &h, int n, long long nnz, std::vector<int> &&rows0, std::vector<int> &&cols0, std::vector<int> &&diag0, int ndeg);
void values_changed(Handle &h);
bool partition_set_values(Handle &h, const double *vals, bool on_device);

}  // namespace b200

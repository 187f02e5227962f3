// Multi-GPU layer: one process per GPU, rows partitioned along the ElmerGrid/METIS partitions.
//
// Replaces the MPI path of SParIterSolver (fem/src/SParIterSolver.F90:123-836 SplitMatrix, 2403-2620 Solve,
// 2630-2745 SParMatrixVector) and SParIterComm (fem/src/SParIterComm.F90:4719-4969 interface exchange,
// 5081-5136 SParDotProd/SParNorm).  The input is what the reference's own GPU bridges already build
// (fem/src/SolverUtils.F90:15461-15579): complete owned rows in continuous global numbering
// (SParIterSolver.F90:1453-1488).  The halo lists follow the canonical form of the reference's only
// x-halo implementation, elmer_distribute_matrix (fem/src/rocalution.cpp:64-372):
//   send list to rank r  = owned rows (ascending) with at least one column owned by r   (121-156)
//   receive list from r  = r's send list to this rank, in r's order                     (222-248)
//   ghost slot of gid g  = position of g in the receive lists concatenated by rank      (286-297)
// Per SpMV: pack kernel -> grouped ncclSend/ncclRecv per neighbour straight into the ghost tail of x
// (message sizes fixed at setup: no per-SpMV size handshake, no barrier) overlapped with the
// owned x owned product; then the ghost block is added.  Per reduction point: one ncclAllReduce of
// the stacked partial sums.  ILU(0) acts on the owned x owned block only (block Jacobi), exactly as
// the reference preconditions InsideMatrix (SParIterSolver.F90:2491-2497).
#include "common.cuh"
#include "kernels.cuh"
#include "krylov.h"
#include "../../include/elmer_b200.h"
#include <nccl.h>      // types and prototypes only: the library is bound at run time (see NcclApi)
#include <dlfcn.h>
#include <time.h>
#include <algorithm>
#include <unordered_map>

namespace b200 {

// NCCL is loaded with dlopen on first use, not linked: single-GPU users need no NCCL at all, and a host
// process that already carries its own libnccl.so.2 (e.g. a Python process with torch imported) shares
// that copy instead of colliding with a second one.  B200_NCCL_LIB overrides the soname.
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
};
static NcclApi &nccl() {
  static NcclApi api;
  static bool loaded = false;
  if (loaded) return api;
  const char *name = getenv("B200_NCCL_LIB");
  void *lib = dlopen((name && *name) ? name : "libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!lib) throw Error(std::string("elmer_b200: cannot load NCCL (") + dlerror() + "); multi-GPU runs need libnccl.so.2");
  auto sym = [&](const char *n) { void *p = dlsym(lib, n); if (!p) throw Error(std::string("elmer_b200: NCCL symbol missing: ") + n); return p; };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  loaded = true;
  return api;
}

#define B200_NCCL(x)                                                                              \
  do {                                                                                            \
    ncclResult_t r__ = (x);                                                                       \
    if (r__ != ncclSuccess) {                                                                     \
      char m__[512];                                                                              \
      snprintf(m__, sizeof m__, "NCCL error '%s' in %s at %s:%d", nccl().GetErrorString(r__), #x,     \
               __FILE__, __LINE__);                                                               \
      throw b200::Error(m__);                                                                     \
    }                                                                                             \
  } while (0)

struct Halo {
  int nneigh = 0, nsend = 0, nghost = 0;
  std::vector<int> neigh, send_ptr, send_idx, recv_ptr, ghost_gid;   // the bit-exact integer deliverables
  DBuf<int> d_send_idx; DBuf<double> d_sendbuf;
  Sell G;                                   // ghost block over the rows that have ghost columns
  DBuf<int> d_gcols, d_gsrc, d_oosrc;       // ghost CRS columns (n_own + slot), value source positions
  DBuf<double> d_vals_in, d_gvals;
  long long nnz_in = 0, nnz_g = 0;
  cudaEvent_t ev_pack = nullptr, ev_comm = nullptr;
  // peer-memory exchange (same node): the pack kernel stores straight into the neighbours' receive buffers
  bool p2p = false;
  void *shm = nullptr;                                  // IPC-shared: recv[2][nghost] f64 | ready[nneigh] u64 | ack[nneigh] u64
  double *recv[2] = {nullptr, nullptr};
  unsigned long long *ready = nullptr, *ack = nullptr;  // ready[q]: last exchange whose data from neighbour q has landed here;
                                                        // ack[q]: last exchange of mine neighbour q has consumed
  std::vector<void *> peer_base;                        // opened IPC mappings, one per neighbour
  DBuf<unsigned char> d_peers; DBuf<unsigned int> d_ticket;
  unsigned long long seq = 0;
};

// what rank `me` needs to know about neighbour q to push into / acknowledge to its memory
struct PeerDesc {
  double *rdata[2];               // my segment of the neighbour's receive buffers
  unsigned long long *rready;     // the neighbour's ready flag for me
  unsigned long long *rack;       // the neighbour's ack flag for me
  int begin, end;                 // my send segment [begin, end)
};


// ---------------------------------------------------------------------------------------------
// Host-side planning (no CUDA, no NCCL): also exported on its own so that the integer lists can be
// checked without a GPU.
//
// Send lists (rocalution.cpp:121-156): owned rows ascending, de-duplicated per destination rank, as
// GLOBAL 0-based row ids.
static void plan_send_lists(int np, int me, int N, const int *rows, const int *cols, int base, const int *goffset, long long gn,
                            std::vector<std::vector<int>> &boundary, long long &nnz_oo, long long &nnz_g) {
  const int lo = goffset[me], hi = goffset[me + 1];
  boundary.assign(np, std::vector<int>());
  std::vector<int> last(np, -1);
  nnz_oo = 0; nnz_g = 0;
  for (int i = 0; i < N; ++i) {
    for (int p = rows[i] - base; p < rows[i + 1] - base; ++p) {
      const int c = cols[p] - base;
      B200_REQUIRE(c >= 0 && c < gn, "global column out of range");
      if (c >= lo && c < hi) { ++nnz_oo; continue; }
      int r = (int)(std::upper_bound(goffset, goffset + np + 1, c) - goffset) - 1;
      if (last[r] != i) { boundary[r].push_back(i + lo); last[r] = i; }
      ++nnz_g;
    }
  }
}

// Split of the complete owned rows into the owned x owned block (local columns, diagonal positions)
// and the ghost block (column = N + ghost slot), given the ghost ids in receive order
// (rocalution.cpp:286-338).  *src arrays give the position of each entry in the caller's value array.
struct SplitPlan { std::vector<int> r0, c0, d0, oosrc, gperm, gstart, glen, gcols, gsrc; };
static void plan_split(int N, const int *rows, const int *cols, int base, int lo, int hi, const std::vector<int> &ghost_gid,
                       long long nnz_oo, long long nnz_g, SplitPlan &S) {
  std::unordered_map<int, int> slot_of; slot_of.reserve(ghost_gid.size() * 2 + 1);
  for (int k = 0; k < (int)ghost_gid.size(); ++k) slot_of[ghost_gid[k]] = k;
  S.r0.assign((size_t)N + 1, 0); S.c0.resize((size_t)nnz_oo); S.d0.assign(N, -1); S.oosrc.resize((size_t)nnz_oo);
  S.gcols.resize((size_t)nnz_g); S.gsrc.resize((size_t)nnz_g); S.gperm.clear(); S.gstart.clear(); S.glen.clear();
  long long l = 0, k = 0;
  for (int i = 0; i < N; ++i) {
    long long k_row = k;
    for (int p = rows[i] - base; p < rows[i + 1] - base; ++p) {
      const int c = cols[p] - base;
      if (c >= lo && c < hi) {
        if (c - lo == i) S.d0[i] = (int)l;
        S.c0[l] = c - lo; S.oosrc[l] = p; ++l;
      } else {
        auto it = slot_of.find(c);
        B200_REQUIRE(it != slot_of.end(), "ghost column not provided by its owner: the sparsity pattern is not structurally symmetric");
        S.gcols[k] = N + it->second; S.gsrc[k] = p; ++k;
      }
    }
    S.r0[i + 1] = (int)l;
    B200_REQUIRE(S.d0[i] >= 0, "owned row without a diagonal entry");
    if (k > k_row) { S.gperm.push_back(i); S.gstart.push_back((int)k_row); S.glen.push_back((int)(k - k_row)); }
  }
}

size_t vec_len(const Handle &h) { return (size_t)h.n + (h.halo ? (size_t)h.halo->nghost : 0); }

// Where rank `me` finds its segment in the receive area of rank r, derived from the send-count matrix cnt[s*np + d]
// (entries rank s sends to rank d) with the rules b200_set_partition applies on rank r itself: r's neighbours are the ranks
// it exchanges anything with, ascending; its receive offsets are the prefix sums of what each of them sends to r.
// qprime = index of `me` in r's neighbour list (-1: not a neighbour), off_me = r.recv_ptr[qprime].
static void peer_layout(int np, int me, int r, const int *cnt, int &qprime, int &nneigh_r, long long &nghost_r, long long &off_me) {
  qprime = -1; nneigh_r = 0; nghost_r = 0; off_me = 0;
  for (int s2 = 0; s2 < np; ++s2) {
    if (s2 == r) continue;
    const int to_r = cnt[(size_t)s2 * np + r], from_r = cnt[(size_t)r * np + s2];
    if (!(to_r || from_r)) continue;
    if (s2 == me) { qprime = nneigh_r; off_me = nghost_r; }
    ++nneigh_r; nghost_r += to_r;
  }
}

static void p2p_release(Handle &h, Halo &H) {
  if (h.stream) cudaStreamSynchronize(h.stream);
  // Neighbours write into this rank's area until they have acknowledged its last exchange (the ack is their last
  // store here): wait for those acks (bounded) before the memory goes away.
  if (H.p2p && H.seq > 0 && H.nneigh > 0) {
    std::vector<unsigned long long> ack(H.nneigh);
    for (int tries = 0; tries < 2000; ++tries) {
      if (cudaMemcpy(ack.data(), H.ack, ack.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); break; }
      bool all = true;
      for (unsigned long long a : ack) all = all && a >= H.seq;
      if (all) break;
      struct timespec ts = {0, 1000000}; nanosleep(&ts, nullptr);
    }
  }
  for (void *b : H.peer_base) if (b) cudaIpcCloseMemHandle(b);
  H.peer_base.clear();
  if (H.shm) cudaFree(H.shm);
  H.shm = nullptr; H.p2p = false; H.d_peers.release(); H.d_ticket.release();
}

// Sets up the peer-memory exchange: every rank publishes the IPC handle of its receive area; the layout of a
// neighbour's area follows from the send-count matrix all ranks already hold.  Collective; all ranks
// end up with the same answer (peer path on every rank, or the NCCL path on every rank).
static void p2p_setup(Handle &h, Halo &H, const std::vector<int> &allcnt) {
  const int np = h.nranks, me = h.rank;
  const char *e = getenv("B200_HALO_P2P");
  int ok = !(e && atoi(e) == 0) && H.nneigh <= 64;
  cudaStream_t st = h.stream;
  const size_t bytes = std::max<size_t>(64, (size_t)2 * H.nghost * sizeof(double) + (size_t)2 * H.nneigh * sizeof(unsigned long long));
  cudaIpcMemHandle_t mine; memset(&mine, 0, sizeof mine);
  if (ok && (cudaMalloc(&H.shm, bytes) != cudaSuccess || cudaMemset(H.shm, 0, bytes) != cudaSuccess ||
             cudaIpcGetMemHandle(&mine, H.shm) != cudaSuccess)) { cudaGetLastError(); ok = 0; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::vector<cudaIpcMemHandle_t> all(np);
  DBuf<char> d_one, d_all; d_one.ensure(64); d_all.ensure((size_t)64 * np);
  B200_CUDA(cudaMemcpyAsync(d_one.p, &mine, 64, cudaMemcpyHostToDevice, st));
  B200_NCCL(nccl().AllGather(d_one.p, d_all.p, 64, ncclChar, (ncclComm_t)h.nccl, st));
  B200_CUDA(cudaMemcpyAsync(all.data(), d_all.p, (size_t)64 * np, cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  std::vector<PeerDesc> peers(std::max(H.nneigh, 1));
  H.peer_base.assign(H.nneigh, nullptr);
  for (int q = 0; q < H.nneigh && ok; ++q) {
    const int r = H.neigh[q];
    int qprime = -1, nneigh_r = 0; long long nghost_r = 0, off_me = 0;
    peer_layout(np, me, r, allcnt.data(), qprime, nneigh_r, nghost_r, off_me);
    if (qprime < 0 || cudaIpcOpenMemHandle(&H.peer_base[q], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); H.peer_base[q] = nullptr; ok = 0; break; }
    double *base = (double *)H.peer_base[q];
    unsigned long long *flags = (unsigned long long *)(base + 2 * nghost_r);
    peers[q].rdata[0] = base + off_me; peers[q].rdata[1] = base + nghost_r + off_me;
    peers[q].rready = flags + qprime; peers[q].rack = flags + nneigh_r + qprime;
    peers[q].begin = H.send_ptr[q]; peers[q].end = H.send_ptr[q + 1];
  }
  // agree over all ranks
  DBuf<int> d_ok; d_ok.ensure(1);
  B200_CUDA(cudaMemcpyAsync(d_ok.p, &ok, sizeof(int), cudaMemcpyHostToDevice, st));
  B200_NCCL(nccl().AllReduce(d_ok.p, d_ok.p, 1, ncclInt32, ncclMin, (ncclComm_t)h.nccl, st));
  B200_CUDA(cudaMemcpyAsync(&ok, d_ok.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  B200_CUDA(cudaStreamSynchronize(st));
  d_one.release(); d_all.release(); d_ok.release();
  if (!ok) { p2p_release(h, H); return; }
  H.recv[0] = (double *)H.shm; H.recv[1] = H.recv[0] + H.nghost;
  H.ready = (unsigned long long *)(H.recv[0] + 2 * (size_t)H.nghost); H.ack = H.ready + H.nneigh;
  H.d_peers.ensure(peers.size() * sizeof(PeerDesc)); H.d_ticket.ensure(2);
  B200_CUDA(cudaMemcpyAsync(H.d_peers.p, peers.data(), peers.size() * sizeof(PeerDesc), cudaMemcpyHostToDevice, st));
  B200_CUDA(cudaMemsetAsync(H.d_ticket.p, 0, 2 * sizeof(unsigned int), st));
  B200_CUDA(cudaStreamSynchronize(st));
  H.seq = 0; H.p2p = true;
}

void halo_release(Handle &h) {
  if (h.halo) {
    Halo &H = *h.halo;
    H.d_send_idx.release(); H.d_sendbuf.release(); H.G.release(); H.d_gcols.release(); H.d_gsrc.release();
    H.d_oosrc.release(); H.d_vals_in.release(); H.d_gvals.release();
    if (H.ev_pack) cudaEventDestroy(H.ev_pack);
    if (H.ev_comm) cudaEventDestroy(H.ev_comm);
    p2p_release(h, H);
    delete h.halo; h.halo = nullptr;
  }
  if (h.nccl) { nccl().CommDestroy((ncclComm_t)h.nccl); h.nccl = nullptr; }
}

void comm_allreduce_sum(Handle &h, double *d, int count) {
  B200_REQUIRE(h.nccl, "reduction over ranks requested without b200_comm_init");
  B200_NCCL(nccl().AllReduce(d, d, count, ncclDouble, ncclSum, (ncclComm_t)h.nccl, h.stream));
  h.st_launch++;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_pack(int nsend, const int *__restrict__ idx, const double *__restrict__ x, double *__restrict__ buf) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nsend; k += gridDim.x * blockDim.x) buf[k] = x[idx[k]];
}
// ---- peer-memory exchange -------------------------------------------------------------------
// Replaces Send_LocIf/Recv_LocIf (SParIterComm.F90:4719-4969) on one node: pack and send are ONE kernel
// that gathers the boundary entries of x and stores them over NVLink into the receive buffer of the
// owning neighbour, then raises that neighbour's `ready` flag (release.sys).  The receiver's ghost-block
// kernel waits on its flags, consumes the buffer and acknowledges into the sender's memory; two
// buffers (exchange parity) + the ack let a rank run at most two products ahead.  No second stream,
// no NCCL kernel competing for SMs with the persistent SpMV grid, no host involvement.
constexpr long long HALO_SPIN_LIMIT = 1LL << 28;
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool last_block(unsigned int *ticket) {
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
    if (last) *ticket = 0u;
  }
  __syncthreads();
  return last;
}
__global__ void __launch_bounds__(256) k_pack_push(int nsend, const int *__restrict__ idx, const double *__restrict__ x, const PeerDesc *__restrict__ peers,
                                                    int nneigh, const unsigned long long *ack, unsigned long long seq, unsigned int *ticket, Ctrl *ctrl, int honor_skip) {
  // (no early exit on ctrl->done == 1: the flag protocol must advance identically on every rank.  done == 2 marks a section every rank
  // skips as a whole -- the value derives from all-reduced scalars, identical on all ranks -- and a skipped exchange leaves ready / ack
  // at their previous sequence numbers, which is exactly what the next exchange expects.)
  if (honor_skip && ctrl->done == 2) return;
  // the buffer of this parity was last filled for exchange seq-2: wait until every neighbour has consumed that one
  if ((int)threadIdx.x < nneigh && seq > 2) {
    long long spins = 0;
    while (ld_acquire_sys(ack + threadIdx.x) + 2 < seq) {
      if (++spins > HALO_SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
      __nanosleep(40);
    }
  }
  __syncthreads();
  const int par = (int)(seq & 1ULL);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nsend; k += gridDim.x * blockDim.x) {
    int q = 0;
    while (k >= peers[q].end) ++q;
    peers[q].rdata[par][k - peers[q].begin] = x[idx[k]];
  }
  __threadfence_system();
  if (last_block(ticket) && (int)threadIdx.x < nneigh) st_release_sys(peers[threadIdx.x].rready, seq);
}
__global__ void __launch_bounds__(256) k_spmv_ghost_p2p(SellView G, const double *xg, double *__restrict__ y, const unsigned long long *ready,
                                                         const PeerDesc *__restrict__ peers, int nneigh, unsigned long long seq, unsigned int *ticket, Ctrl *ctrl, int honor_skip) {
  if (honor_skip && ctrl->done == 2) return;
  if ((int)threadIdx.x < nneigh) {
    long long spins = 0;
    while (ld_acquire_sys(ready + threadIdx.x) < seq) {
      if (++spins > HALO_SPIN_LIMIT) { ctrl->spin_timeout = 1; break; }
      __nanosleep(40);
    }
  }
  __syncthreads();
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = slot < G.nslots ? G.perm[slot] : -1;
  if (row >= 0) {
    const int lane = threadIdx.x & 31;
    const long long p0 = G.ptr[slot >> 5];
    const int len = G.len[slot];
    double acc = 0.0;
    for (int j = 0; j < len; ++j) acc = nfma(acc, __ldcg(xg + G.cols[p0 + j * 32 + lane]), G.vals[p0 + j * 32 + lane]);
    y[row] = __dadd_rn(y[row], acc);
  }
  if (last_block(ticket) && (int)threadIdx.x < nneigh) st_release_sys(peers[threadIdx.x].rack, seq);
}

// y[row] += sum_j G_ij x[n_own + slot_j]   (the product the reference receives as partial sums, SParIterComm.F90:4945-4953)
__global__ void __launch_bounds__(256) k_spmv_ghost(SellView G, const double *__restrict__ x, double *__restrict__ y) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= G.nslots) return;
  const int row = G.perm[slot];
  if (row < 0) return;
  const int lane = threadIdx.x & 31;
  const long long p0 = G.ptr[slot >> 5];
  const int len = G.len[slot];
  double acc = 0.0;
  for (int j = 0; j < len; ++j) acc = nfma(acc, x[G.cols[p0 + j * 32 + lane]], G.vals[p0 + j * 32 + lane]);
  y[row] = __dadd_rn(y[row], acc);
}
__global__ void k_gather_vals(long long n, const int *__restrict__ src, const double *__restrict__ in, double *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[src[i]];
}

void matvec_full(Handle &h, const double *x, double *y) {
  if (!h.halo || h.nranks == 1 || h.halo->nneigh == 0) {
    SpmvArgs a; a.x = x; a.y = y; spmv_launch(h, a, EPI_NONE);
    return;
  }
  Halo &H = *h.halo;
  ncclComm_t comm = (ncclComm_t)h.nccl;
  double *xg = const_cast<double *>(x) + h.n;                      // ghost tail of the operand
  if (H.p2p) {
    const unsigned long long seq = ++H.seq;
    const PeerDesc *peers = (const PeerDesc *)H.d_peers.p;
    k_pack_push<<<std::max(1, std::min((H.nsend + 255) / 256, NUM_SMS * 4)), 256, 0, h.stream>>>(H.nsend, H.d_send_idx.p, x, peers, H.nneigh, H.ack, seq,
                                                                                                 H.d_ticket.p, h.ctrl.p, h.mv_honor_skip ? 1 : 0);
    { SpmvArgs a; a.x = x; a.y = y; if (h.mv_honor_skip) a.ctrl = h.ctrl.p; spmv_launch(h, a, EPI_NONE); }   // owned x owned while the neighbours' entries arrive
    k_spmv_ghost_p2p<<<std::max(1, (H.G.nslots + 255) / 256), 256, 0, h.stream>>>(H.G.view(), H.recv[seq & 1ULL] - h.n, y, H.ready, peers, H.nneigh, seq,
                                                                                   H.d_ticket.p + 1, h.ctrl.p, h.mv_honor_skip ? 1 : 0);
    B200_CUDA(cudaGetLastError());
    h.st_launch += 3;
    return;
  }
  if (H.nsend) k_pack<<<std::max(1, std::min((H.nsend + 255) / 256, NUM_SMS * 4)), 256, 0, h.stream>>>(H.nsend, H.d_send_idx.p, x, H.d_sendbuf.p);
  B200_CUDA(cudaEventRecord(H.ev_pack, h.stream));
  B200_CUDA(cudaStreamWaitEvent(h.stream2, H.ev_pack, 0));
  B200_NCCL(nccl().GroupStart());
  for (int q = 0; q < H.nneigh; ++q) {
    int ns = H.send_ptr[q + 1] - H.send_ptr[q], nr = H.recv_ptr[q + 1] - H.recv_ptr[q];
    if (ns) B200_NCCL(nccl().Send(H.d_sendbuf.p + H.send_ptr[q], ns, ncclDouble, H.neigh[q], comm, h.stream2));
    if (nr) B200_NCCL(nccl().Recv(xg + H.recv_ptr[q], nr, ncclDouble, H.neigh[q], comm, h.stream2));
  }
  B200_NCCL(nccl().GroupEnd());
  B200_CUDA(cudaEventRecord(H.ev_comm, h.stream2));
  { SpmvArgs a; a.x = x; a.y = y; spmv_launch(h, a, EPI_NONE); }   // owned x owned, overlaps the exchange
  B200_CUDA(cudaStreamWaitEvent(h.stream, H.ev_comm, 0));
  if (H.G.nslots) k_spmv_ghost<<<(H.G.nslots + 255) / 256, 256, 0, h.stream>>>(H.G.view(), x, y);
  B200_CUDA(cudaGetLastError());
  h.st_launch += 3;
}

// the SpMV epilogues as a separate pass over y (partitioned runs: y is complete only after the ghost block)
template <int EPI>
__global__ void __launch_bounds__(256) k_epi_pass(int n, SpmvArgs a, const double *__restrict__ yin) {
  if (a.ctrl && a.ctrl->done) return;
  constexpr int NV = (EPI == EPI_DOT2) ? 2 : 1;
  double red[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) red[k] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double r = yin[i];
    if (EPI == EPI_DOT1) red[0] += r * a.w[i];
    else if (EPI == EPI_DOT2) { red[0] += r * a.w[i]; red[NV - 1] += r * r; }
    else if (EPI == EPI_RESID) { double d = __dsub_rn(r, a.b[i]); red[0] += d * d; }
    else if (EPI == EPI_BMINUS) { double d = __dsub_rn(a.b[i], r); a.y[i] = d; if (a.y2) a.y2[i] = d; red[0] += d * d; }
  }
  double *out = a.out;
  grid_reduce<NV>(red, a.partials, a.counter, [out](double(&t)[NV]) {
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = t[k];
  });
}

// y = A x with the requested epilogue, single GPU (fused) or partitioned (halo + separate pass)
void spmv_any(Handle &h, SpmvArgs a, int epi) {
  if (!h.halo || h.nranks == 1) { spmv_launch(h, a, epi); return; }
  double *y = a.y;
  if (epi == EPI_RESID) { h.d_tmp.ensure(h.n); y = h.d_tmp.p; }
  matvec_full(h, a.x, y);
  if (epi == EPI_NONE) return;
  a.partials = h.red_partials.p; a.counter = h.red_counters.p;
  int blocks = std::max(1, std::min((h.n + 255) / 256, h.blas_blocks));
  switch (epi) {
    case EPI_DOT1: k_epi_pass<EPI_DOT1><<<blocks, 256, 0, h.stream>>>(h.n, a, y); break;
    case EPI_DOT2: k_epi_pass<EPI_DOT2><<<blocks, 256, 0, h.stream>>>(h.n, a, y); break;
    case EPI_RESID: k_epi_pass<EPI_RESID><<<blocks, 256, 0, h.stream>>>(h.n, a, y); break;
    case EPI_BMINUS: k_epi_pass<EPI_BMINUS><<<blocks, 256, 0, h.stream>>>(h.n, a, y); break;
  }
  B200_CUDA(cudaGetLastError());
  h.st_launch++;
}

bool partition_set_values(Handle &h, const double *vals, bool on_device) {
  Halo &H = *h.halo;
  cudaStream_t st = h.stream;
  H.d_vals_in.ensure(H.nnz_in);
  if (H.nnz_in) {
    B200_CUDA(cudaMemcpyAsync(H.d_vals_in.p, vals, (size_t)H.nnz_in * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (!on_device) h.st_h2d += H.nnz_in * sizeof(double);
  }
  h.d_vals.ensure(h.nnz);
  int gb = NUM_SMS * 8;
  if (h.nnz) k_gather_vals<<<gb, 256, 0, st>>>(h.nnz, H.d_oosrc.p, H.d_vals_in.p, h.d_vals.p);
  H.d_gvals.ensure(H.nnz_g);
  if (H.nnz_g) k_gather_vals<<<gb, 256, 0, st>>>(H.nnz_g, H.d_gsrc.p, H.d_vals_in.p, H.d_gvals.p);
  B200_CUDA(cudaGetLastError());
  h.have_prec = false;
  values_changed(h);
  sell_refresh_values(h, H.G, H.d_gvals.p);
  return true;
}

}  // namespace b200

using namespace b200;

template <class F> static int guarded_c(F f) {
  try { f(); return 0; }
  catch (const std::exception &e) { set_last_error(e.what()); fprintf(stderr, "[elmer_b200] %s\n", e.what()); return 1; }
}

extern "C" {

int b200_comm_unique_id(char *id128) {
  return guarded_c([&] {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id; B200_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, 128);
  });
}

int b200_comm_init(void **handle, const int *nranks, const int *rank, const char *id128) {
  return guarded_c([&] {
    B200_REQUIRE(handle && *handle, "null handle");
    Handle &h = *static_cast<Handle *>(*handle);
    B200_CUDA(cudaSetDevice(h.device));
    B200_REQUIRE(*nranks >= 1 && *rank >= 0 && *rank < *nranks, "bad rank/nranks");
    h.nranks = *nranks; h.rank = *rank;
    if (*nranks == 1) return;
    ncclUniqueId id; memcpy(&id, id128, 128);
    ncclComm_t comm;
    B200_NCCL(nccl().CommInitRank(&comm, *nranks, id, *rank));
    h.nccl = comm;
  });
}

int b200_set_partition(void **handle, const int *gn, const int *n_own, const int *nnz, const int *rows,
                       const int *cols, const int *goffset, const int *index_base, const int *ndeg) {
  return guarded_c([&] {
    B200_REQUIRE(handle && *handle, "null handle");
    Handle &h = *static_cast<Handle *>(*handle);
    B200_CUDA(cudaSetDevice(h.device));
    const int np = h.nranks, me = h.rank, N = *n_own, base = *index_base;
    const long long NNZ = *nnz;
    B200_REQUIRE(np == 1 || h.nccl, "b200_set_partition before b200_comm_init");
    B200_REQUIRE(base == 0 || base == 1, "index_base must be 0 or 1");
    B200_REQUIRE(goffset[0] == 0 && goffset[np] == *gn, "goffset must run from 0 to gn");
    const int lo = goffset[me], hi = goffset[me + 1];
    B200_REQUIRE(hi - lo == N, "n_own inconsistent with goffset");
    // per-rank sizes are int32 by the ABI (as Elmer's Matrix_t): the owned rows and THEIR entries must fit, the global system need not
    // (C3: 2.0e9 global entries, 5.0e8 per rank on 4 ranks); a wrapped-around row pointer is caught here
    B200_REQUIRE(NNZ >= 0 && (long long)rows[N] - rows[0] == NNZ && NNZ < 2147483647LL, "b200_set_partition: this rank's entries do not fit int32 row pointers (nnz inconsistent with rows)");
    h.gn = *gn; h.index_base = base;
    if (h.halo) { Halo *old = h.halo; h.halo = nullptr; p2p_release(h, *old); old->G.release(); delete old; }
    Halo *Hp = new Halo(); h.halo = Hp; Halo &H = *Hp;
    B200_CUDA(cudaEventCreateWithFlags(&H.ev_pack, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&H.ev_comm, cudaEventDisableTiming));
    H.nnz_in = NNZ;

    std::vector<std::vector<int>> boundary;
    long long nnz_oo = 0, nnz_g = 0;
    plan_send_lists(np, me, N, rows, cols, base, goffset, *gn, boundary, nnz_oo, nnz_g);
    // ---- exchange list sizes and lists
    std::vector<int> sendcnt(np, 0), allcnt((size_t)np * np, 0);
    for (int r = 0; r < np; ++r) sendcnt[r] = (int)boundary[r].size();
    cudaStream_t st = h.stream;
    if (np > 1) {
      DBuf<int> d_cnt, d_all; d_cnt.ensure(np); d_all.ensure((size_t)np * np);
      B200_CUDA(cudaMemcpyAsync(d_cnt.p, sendcnt.data(), np * sizeof(int), cudaMemcpyHostToDevice, st));
      B200_NCCL(nccl().AllGather(d_cnt.p, d_all.p, np, ncclInt32, (ncclComm_t)h.nccl, st));
      B200_CUDA(cudaMemcpyAsync(allcnt.data(), d_all.p, (size_t)np * np * sizeof(int), cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaStreamSynchronize(st));
      d_cnt.release(); d_all.release();
    }
    H.neigh.clear(); H.send_ptr.assign(1, 0); H.recv_ptr.assign(1, 0);
    for (int r = 0; r < np; ++r) {
      if (r == me) continue;
      int ns = sendcnt[r], nr = allcnt[(size_t)r * np + me];
      if (ns || nr) { H.neigh.push_back(r); H.send_ptr.push_back(H.send_ptr.back() + ns); H.recv_ptr.push_back(H.recv_ptr.back() + nr); }
    }
    H.nneigh = (int)H.neigh.size(); H.nsend = H.send_ptr.back(); H.nghost = H.recv_ptr.back();
    std::vector<int> send_gid(H.nsend); H.send_idx.resize(H.nsend); H.ghost_gid.assign(H.nghost, 0);
    for (int q = 0, k = 0; q < H.nneigh; ++q) for (int g : boundary[H.neigh[q]]) { send_gid[k] = g; H.send_idx[k] = g - lo; ++k; }
    if (np > 1 && (H.nsend || H.nghost)) {
      DBuf<int> d_s, d_r; d_s.ensure(H.nsend); d_r.ensure(H.nghost);
      if (H.nsend) B200_CUDA(cudaMemcpyAsync(d_s.p, send_gid.data(), H.nsend * sizeof(int), cudaMemcpyHostToDevice, st));
      B200_NCCL(nccl().GroupStart());
      for (int q = 0; q < H.nneigh; ++q) {
        int ns = H.send_ptr[q + 1] - H.send_ptr[q], nr = H.recv_ptr[q + 1] - H.recv_ptr[q];
        if (ns) B200_NCCL(nccl().Send(d_s.p + H.send_ptr[q], ns, ncclInt32, H.neigh[q], (ncclComm_t)h.nccl, st));
        if (nr) B200_NCCL(nccl().Recv(d_r.p + H.recv_ptr[q], nr, ncclInt32, H.neigh[q], (ncclComm_t)h.nccl, st));
      }
      B200_NCCL(nccl().GroupEnd());
      if (H.nghost) B200_CUDA(cudaMemcpyAsync(H.ghost_gid.data(), d_r.p, H.nghost * sizeof(int), cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaStreamSynchronize(st));
      d_s.release(); d_r.release();
    }
    SplitPlan sp;
    plan_split(N, rows, cols, base, lo, hi, H.ghost_gid, nnz_oo, nnz_g, sp);
    std::vector<int> &r0 = sp.r0, &c0 = sp.c0, &d0 = sp.d0, &oosrc = sp.oosrc, &gperm = sp.gperm, &gstart = sp.gstart, &glen = sp.glen,
                     &gcols = sp.gcols, &gsrc = sp.gsrc;
    H.nnz_g = nnz_g;
    install_structure(h, N, nnz_oo, std::move(r0), std::move(c0), std::move(d0), ndeg ? *ndeg : 1);
    // ghost block as a SELL over the boundary rows
    int nb = (int)gperm.size(), nslots = ((nb + 31) / 32) * 32;
    gperm.resize(nslots, -1); gstart.resize(nslots, 0); glen.resize(nslots, 0);
    H.G.perm.ensure(nslots); H.G.start.ensure(nslots); H.G.len.ensure(nslots);
    H.d_gcols.ensure(nnz_g); H.d_gsrc.ensure(nnz_g); H.d_oosrc.ensure(nnz_oo); H.d_send_idx.ensure(H.nsend); H.d_sendbuf.ensure(H.nsend);
    if (nslots) {
      B200_CUDA(cudaMemcpyAsync(H.G.perm.p, gperm.data(), nslots * sizeof(int), cudaMemcpyHostToDevice, st));
      B200_CUDA(cudaMemcpyAsync(H.G.start.p, gstart.data(), nslots * sizeof(int), cudaMemcpyHostToDevice, st));
      B200_CUDA(cudaMemcpyAsync(H.G.len.p, glen.data(), nslots * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (nnz_g) {
      B200_CUDA(cudaMemcpyAsync(H.d_gcols.p, gcols.data(), (size_t)nnz_g * sizeof(int), cudaMemcpyHostToDevice, st));
      B200_CUDA(cudaMemcpyAsync(H.d_gsrc.p, gsrc.data(), (size_t)nnz_g * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    if (nnz_oo) B200_CUDA(cudaMemcpyAsync(H.d_oosrc.p, oosrc.data(), (size_t)nnz_oo * sizeof(int), cudaMemcpyHostToDevice, st));
    if (H.nsend) B200_CUDA(cudaMemcpyAsync(H.d_send_idx.p, H.send_idx.data(), H.nsend * sizeof(int), cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaStreamSynchronize(st));
    sell_finish(h, H.G, nslots, true, H.d_gcols.p);
    B200_CUDA(cudaStreamSynchronize(st));
    if (np > 1) p2p_setup(h, H, allcnt);
  });
}

// Host-only planning entry points (usable without a GPU; the caller does the exchange itself).
int b200_partition_send_lists(const int *gn, const int *n_own, const int *rows, const int *cols, const int *goffset,
                              const int *index_base, const int *nranks, const int *rank, int *send_count, int *send_gid) {
  return guarded_c([&] {
    std::vector<std::vector<int>> boundary; long long a = 0, b = 0;
    plan_send_lists(*nranks, *rank, *n_own, rows, cols, *index_base, goffset, *gn, boundary, a, b);
    int k = 0;
    for (int r = 0; r < *nranks; ++r) {
      send_count[r] = (int)boundary[r].size();
      if (send_gid) for (int g : boundary[r]) send_gid[k++] = g;
    }
  });
}

int b200_partition_split(const int *n_own, const int *rows, const int *cols, const int *index_base, const int *lo, const int *hi,
                         const int *nghost, const int *ghost_gid, int *sizes, int *oo_rows, int *oo_cols, int *oo_diag,
                         int *g_rows, int *g_cols) {
  return guarded_c([&] {
    const int N = *n_own, base = *index_base;
    long long nnz_oo = 0, nnz_g = 0;
    for (int i = 0; i < N; ++i) for (int p = rows[i] - base; p < rows[i + 1] - base; ++p) {
      const int c = cols[p] - base; if (c >= *lo && c < *hi) ++nnz_oo; else ++nnz_g;
    }
    sizes[0] = (int)nnz_oo; sizes[1] = (int)nnz_g;
    if (!oo_rows) return;
    std::vector<int> gg(ghost_gid, ghost_gid + *nghost);
    SplitPlan S; plan_split(N, rows, cols, base, *lo, *hi, gg, nnz_oo, nnz_g, S);
    std::copy(S.r0.begin(), S.r0.end(), oo_rows); std::copy(S.c0.begin(), S.c0.end(), oo_cols); std::copy(S.d0.begin(), S.d0.end(), oo_diag);
    // ghost block as CRS over all owned rows
    std::vector<int> gr((size_t)N + 1, 0);
    for (size_t q = 0; q < S.gperm.size(); ++q) gr[S.gperm[q] + 1] = S.glen[q];
    for (int i = 0; i < N; ++i) gr[i + 1] += gr[i];
    std::copy(gr.begin(), gr.end(), g_rows); std::copy(S.gcols.begin(), S.gcols.end(), g_cols);
  });
}

/* host-only: layout of rank r's receive area as rank `me` derives it (out[0] = index of me among r's neighbours or -1,
 * out[1] = r's neighbour count, out[2] = r's ghost count, out[3] = offset of me's segment); cnt = np x np send counts */
int b200_partition_peer_layout(const int *nranks, const int *me, const int *r, const int *cnt, long long *out) {
  return guarded_c([&] {
    int q, nn; long long ng, off;
    peer_layout(*nranks, *me, *r, cnt, q, nn, ng, off);
    out[0] = q; out[1] = nn; out[2] = ng; out[3] = off;
  });
}

int b200_get_halo_plan(void **handle, int *sizes, int *neigh, int *send_ptr, int *send_idx, int *recv_ptr, int *ghost_gid) {
  return guarded_c([&] {
    B200_REQUIRE(handle && *handle, "null handle");
    Handle &h = *static_cast<Handle *>(*handle);
    B200_REQUIRE(h.halo, "no partition set");
    Halo &H = *h.halo;
    sizes[0] = H.nneigh; sizes[1] = H.nsend; sizes[2] = H.nghost;
    if (neigh) std::copy(H.neigh.begin(), H.neigh.end(), neigh);
    if (send_ptr) std::copy(H.send_ptr.begin(), H.send_ptr.end(), send_ptr);
    if (send_idx) std::copy(H.send_idx.begin(), H.send_idx.end(), send_idx);
    if (recv_ptr) std::copy(H.recv_ptr.begin(), H.recv_ptr.end(), recv_ptr);
    if (ghost_gid) std::copy(H.ghost_gid.begin(), H.ghost_gid.end(), ghost_gid);
  });
}

}  // extern "C"

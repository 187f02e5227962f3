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
#include <nccl.h>
#include <algorithm>
#include <unordered_map>

namespace b200 {

#define B200_NCCL(x)                                                                              \
  do {                                                                                            \
    ncclResult_t r__ = (x);                                                                       \
    if (r__ != ncclSuccess) {                                                                     \
      char m__[512];                                                                              \
      snprintf(m__, sizeof m__, "NCCL error '%s' in %s at %s:%d", ncclGetErrorString(r__), #x,     \
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
};

size_t vec_len(const Handle &h) { return (size_t)h.n + (h.halo ? (size_t)h.halo->nghost : 0); }

void halo_release(Handle &h) {
  if (h.halo) {
    Halo &H = *h.halo;
    H.d_send_idx.release(); H.d_sendbuf.release(); H.G.release(); H.d_gcols.release(); H.d_gsrc.release();
    H.d_oosrc.release(); H.d_vals_in.release(); H.d_gvals.release();
    if (H.ev_pack) cudaEventDestroy(H.ev_pack);
    if (H.ev_comm) cudaEventDestroy(H.ev_comm);
    delete h.halo; h.halo = nullptr;
  }
  if (h.nccl) { ncclCommDestroy((ncclComm_t)h.nccl); h.nccl = nullptr; }
}

void comm_allreduce_sum(Handle &h, double *d, int count) {
  B200_REQUIRE(h.nccl, "reduction over ranks requested without b200_comm_init");
  B200_NCCL(ncclAllReduce(d, d, count, ncclDouble, ncclSum, (ncclComm_t)h.nccl, h.stream));
  h.st_launch++;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_pack(int nsend, const int *__restrict__ idx, const double *__restrict__ x, double *__restrict__ buf) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nsend; k += gridDim.x * blockDim.x) buf[k] = x[idx[k]];
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
  if (H.nsend) k_pack<<<std::max(1, std::min((H.nsend + 255) / 256, NUM_SMS * 4)), 256, 0, h.stream>>>(H.nsend, H.d_send_idx.p, x, H.d_sendbuf.p);
  B200_CUDA(cudaEventRecord(H.ev_pack, h.stream));
  B200_CUDA(cudaStreamWaitEvent(h.stream2, H.ev_pack, 0));
  B200_NCCL(ncclGroupStart());
  for (int q = 0; q < H.nneigh; ++q) {
    int ns = H.send_ptr[q + 1] - H.send_ptr[q], nr = H.recv_ptr[q + 1] - H.recv_ptr[q];
    if (ns) B200_NCCL(ncclSend(H.d_sendbuf.p + H.send_ptr[q], ns, ncclDouble, H.neigh[q], comm, h.stream2));
    if (nr) B200_NCCL(ncclRecv(xg + H.recv_ptr[q], nr, ncclDouble, H.neigh[q], comm, h.stream2));
  }
  B200_NCCL(ncclGroupEnd());
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
    ncclUniqueId id; B200_NCCL(ncclGetUniqueId(&id));
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
    B200_NCCL(ncclCommInitRank(&comm, *nranks, id, *rank));
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
    h.gn = *gn; h.index_base = base;
    if (h.halo) { cudaStream_t s = h.stream; (void)s; Halo *old = h.halo; h.halo = nullptr; old->G.release(); delete old; }
    Halo *Hp = new Halo(); h.halo = Hp; Halo &H = *Hp;
    B200_CUDA(cudaEventCreateWithFlags(&H.ev_pack, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&H.ev_comm, cudaEventDisableTiming));
    H.nnz_in = NNZ;

    // ---- send lists (rocalution.cpp:121-156): rows ascending, de-duplicated per destination rank
    std::vector<std::vector<int>> boundary(np);
    std::vector<int> last(np, -1);
    long long nnz_oo = 0, nnz_g = 0;
    for (int i = 0; i < N; ++i) {
      for (int p = rows[i] - base; p < rows[i + 1] - base; ++p) {
        const int c = cols[p] - base;
        B200_REQUIRE(c >= 0 && c < *gn, "global column out of range");
        if (c >= lo && c < hi) { ++nnz_oo; continue; }
        int r = (int)(std::upper_bound(goffset, goffset + np + 1, c) - goffset) - 1;
        if (last[r] != i) { boundary[r].push_back(i + lo); last[r] = i; }
        ++nnz_g;
      }
    }
    // ---- exchange list sizes and lists
    std::vector<int> sendcnt(np, 0), allcnt((size_t)np * np, 0);
    for (int r = 0; r < np; ++r) sendcnt[r] = (int)boundary[r].size();
    cudaStream_t st = h.stream;
    if (np > 1) {
      DBuf<int> d_cnt, d_all; d_cnt.ensure(np); d_all.ensure((size_t)np * np);
      B200_CUDA(cudaMemcpyAsync(d_cnt.p, sendcnt.data(), np * sizeof(int), cudaMemcpyHostToDevice, st));
      B200_NCCL(ncclAllGather(d_cnt.p, d_all.p, np, ncclInt32, (ncclComm_t)h.nccl, st));
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
      B200_NCCL(ncclGroupStart());
      for (int q = 0; q < H.nneigh; ++q) {
        int ns = H.send_ptr[q + 1] - H.send_ptr[q], nr = H.recv_ptr[q + 1] - H.recv_ptr[q];
        if (ns) B200_NCCL(ncclSend(d_s.p + H.send_ptr[q], ns, ncclInt32, H.neigh[q], (ncclComm_t)h.nccl, st));
        if (nr) B200_NCCL(ncclRecv(d_r.p + H.recv_ptr[q], nr, ncclInt32, H.neigh[q], (ncclComm_t)h.nccl, st));
      }
      B200_NCCL(ncclGroupEnd());
      if (H.nghost) B200_CUDA(cudaMemcpyAsync(H.ghost_gid.data(), d_r.p, H.nghost * sizeof(int), cudaMemcpyDeviceToHost, st));
      B200_CUDA(cudaStreamSynchronize(st));
      d_s.release(); d_r.release();
    }
    // ---- ghost slot map (rocalution.cpp:286-297)
    std::unordered_map<int, int> slot_of; slot_of.reserve((size_t)H.nghost * 2);
    for (int k = 0; k < H.nghost; ++k) slot_of[H.ghost_gid[k]] = k;
    // ---- split into the owned x owned block (local columns) and the ghost block
    std::vector<int> r0((size_t)N + 1, 0), c0((size_t)nnz_oo), d0(N, -1), oosrc((size_t)nnz_oo);
    std::vector<int> gperm, gstart, glen, gcols((size_t)nnz_g), gsrc((size_t)nnz_g);
    long long l = 0, k = 0;
    for (int i = 0; i < N; ++i) {
      long long k_row = k;
      for (int p = rows[i] - base; p < rows[i + 1] - base; ++p) {
        const int c = cols[p] - base;
        if (c >= lo && c < hi) {
          if (c - lo == i) d0[i] = (int)l;
          c0[l] = c - lo; oosrc[l] = p; ++l;
        } else {
          auto it = slot_of.find(c);
          B200_REQUIRE(it != slot_of.end(), "ghost column not provided by its owner: the sparsity pattern is not structurally symmetric");
          gcols[k] = N + it->second; gsrc[k] = p; ++k;
        }
      }
      r0[i + 1] = (int)l;
      B200_REQUIRE(d0[i] >= 0, "owned row without a diagonal entry");
      if (k > k_row) { gperm.push_back(i); gstart.push_back((int)k_row); glen.push_back((int)(k - k_row)); }
    }
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

// fp64 SpMV on the SELL-32 operand, with the Krylov methods' dot/norm work fused into its epilogue.
//
// Replaces CRS_MatrixVectorProd (fem/src/CRSMatrix.F90:4744-4905) and its twin CRS_MatrixVectorMultiply
// (1496-1620).  One thread owns one row and walks it left to right, so for ndeg = 1 the products are
// added in exactly the reference's order (4859-4866) with separate multiply and add roundings (the
// reference is compiled without FMA contraction): the result is bit-identical to the CPU loop.  For
// ndeg in {2,3,4,5,6,8,10} the NACC partial sums r1..r5 of 4794-4856 are kept and added left to right.
//
// Memory behaviour: slice storage is column-major within 32 rows, so every warp load of values (256 B)
// and column ids (128 B) is one fully coalesced, L1-bypassing streaming request; the x gathers of 32
// neighbouring rows hit neighbouring addresses and are served by L1/L2 (x is re-read ~27 times and
// stays L2 resident: 65 MB at 8 M rows vs 126 MB of L2).  Persistent grid sized in multiples of the
// 148 SMs, grid-stride over slices.
#include "common.cuh"
#include "kernels.cuh"

namespace b200 {

// BLK: every group of NACC entries is a run of consecutive columns (checked at set_structure): one index load per group,
// operands at c, c+1, ..., as the reference's ndeg loops address them; two groups in flight.
template <int NACC, int EPI, bool BLK>
__global__ void __launch_bounds__(256) k_spmv_sell(SellView A, int n, SpmvArgs a) {
  if (a.ctrl && a.ctrl->done) return;
  constexpr int U = (NACC == 1) ? 4 : ((NACC == 2) ? 4 : (BLK ? 2 * NACC : NACC));   // loads in flight per thread
  constexpr int NV = (EPI == EPI_NONE) ? 1 : ((EPI == EPI_DOT2) ? 2 : 1);
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double red[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) red[k] = 0.0;
  const double *__restrict__ x = a.x;

  for (int slice = gwarp; slice < A.nslices; slice += nwarps) {
    const long long p0 = A.ptr[slice];
    const int W = (int)((A.ptr[slice + 1] - p0) >> 5);
    const int row = slice * 32 + lane;
    const int len = A.len[row];
    const int *__restrict__ cp = A.cols + p0 + lane;
    const double *__restrict__ vp = A.vals + p0 + lane;
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    int j = 0;
    for (; j + U <= W; j += U) {
      int c[U]; double v[U], xv[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (!BLK || k % NACC == 0) c[k] = ld_stream(cp + (j + k) * 32); else c[k] = c[k - k % NACC] + k % NACC;
        v[k] = ld_stream(vp + (j + k) * 32);
      }
#pragma unroll
      for (int k = 0; k < U; ++k) xv[k] = __ldg(x + c[k]);
#pragma unroll
      for (int k = 0; k < U; ++k) if (j + k < len) acc[k % NACC] = nfma(acc[k % NACC], xv[k], v[k]);
    }
    for (; j < W; j += NACC) {            // tail: W is a multiple of NACC (checked at set_structure)
      int c0 = 0;
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        int c = (!BLK || k == 0) ? ld_stream(cp + (j + k) * 32) : c0 + k; double v = ld_stream(vp + (j + k) * 32);
        if (k == 0) c0 = c;
        double xv = __ldg(x + c);
        if (j + k < len) acc[k] = nfma(acc[k], xv, v);
      }
    }
    double r = acc[0];
#pragma unroll
    for (int k = 1; k < NACC; ++k) r = __dadd_rn(r, acc[k]);

    if (row < n) {
      if (EPI == EPI_NONE) {
        a.y[row] = r;
      } else if (EPI == EPI_DOT1) {          // y = A x ; sum y*w
        a.y[row] = r;
        red[0] += r * a.w[row];
      } else if (EPI == EPI_DOT2) {          // y = A x ; sum y*w , sum y*y
        a.y[row] = r;
        red[0] += r * a.w[row];
        red[NV - 1] += r * r;
      } else if (EPI == EPI_RESID) {         // sum (A x - b)^2, nothing stored
        double d = __dsub_rn(r, a.b[row]);
        red[0] += d * d;
      } else if (EPI == EPI_BMINUS) {        // y = b - A x (and y2 = y) ; sum y*y
        double d = __dsub_rn(a.b[row], r);
        a.y[row] = d;
        if (a.y2) a.y2[row] = d;
        red[0] += d * d;
      }
    }
  }
  if (EPI != EPI_NONE) {
    double *out = a.out;
    grid_reduce<NV>(red, a.partials, a.counter, [out](double(&t)[NV]) {
#pragma unroll
      for (int k = 0; k < NV; ++k) out[k] = t[k];
    });
  }
}

template <int NACC, bool BLK>
static void launch_epi2(Handle &h, int blocks, const SellView &A, const SpmvArgs &a, int epi) {
  cudaStream_t st = h.stream;
  switch (epi) {
    case EPI_NONE:   k_spmv_sell<NACC, EPI_NONE, BLK><<<blocks, 256, 0, st>>>(A, h.n, a); break;
    case EPI_DOT1:   k_spmv_sell<NACC, EPI_DOT1, BLK><<<blocks, 256, 0, st>>>(A, h.n, a); break;
    case EPI_DOT2:   k_spmv_sell<NACC, EPI_DOT2, BLK><<<blocks, 256, 0, st>>>(A, h.n, a); break;
    case EPI_RESID:  k_spmv_sell<NACC, EPI_RESID, BLK><<<blocks, 256, 0, st>>>(A, h.n, a); break;
    case EPI_BMINUS: k_spmv_sell<NACC, EPI_BMINUS, BLK><<<blocks, 256, 0, st>>>(A, h.n, a); break;
    default: throw Error("spmv: bad epilogue");
  }
}
template <int NACC>
static void launch_epi(Handle &h, int blocks, const SellView &A, const SpmvArgs &a, int epi) {
  if (NACC > 1 && h.nacc_blocked) launch_epi2<NACC, true>(h, blocks, A, a, epi);
  else launch_epi2<NACC, false>(h, blocks, A, a, epi);
}

void spmv_launch(Handle &h, SpmvArgs a, int epi) {
  if (h.n == 0) return;
  SellView A = h.A.view();
  int blocks = h.spmv_blocks > 0 ? h.spmv_blocks : NUM_SMS * 8;
  int need = (A.nslices + 7) / 8;
  if (blocks > need) blocks = need;
  if (blocks > MAX_RED_BLOCKS) blocks = MAX_RED_BLOCKS;
  if (!a.partials) { a.partials = h.red_partials.p; a.counter = h.red_counters.p; }
  switch (h.nacc) {
    case 1: launch_epi<1>(h, blocks, A, a, epi); break;
    case 2: launch_epi<2>(h, blocks, A, a, epi); break;
    case 3: launch_epi<3>(h, blocks, A, a, epi); break;
    case 4: launch_epi<4>(h, blocks, A, a, epi); break;
    case 5: launch_epi<5>(h, blocks, A, a, epi); break;
    default: throw Error("spmv: bad ndeg accumulator count");
  }
  B200_CUDA(cudaGetLastError());
  h.st_matvec++; h.st_launch++;
}

}  // namespace b200

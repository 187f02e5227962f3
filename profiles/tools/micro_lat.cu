// Micro-benchmarks behind the triangular-solve design: dependent-issue latency of the fp64 pipe, shuffle and
// shared-memory round trips of a single warp, the L2 round trip of a relaxed load, and the store -> poll hand-off between two SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o micro_lat micro_lat.cu && ./micro_lat
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void k_chain(double *out, long long *cyc, double a, double b, int warps) {
  const int N = 4096;
  double x = a + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __dadd_rn(x, b);
  long long t1 = clock64();
  double y = a + threadIdx.x;
#pragma unroll 16
  for (int i = 0; i < N; ++i) y = __dmul_rn(y, b);
  long long t2 = clock64();
  double z = a;
#pragma unroll 16
  for (int i = 0; i < N; ++i) z = __dsub_rn(z, __dmul_rn(b, z));
  long long t3 = clock64();
  double w = a + threadIdx.x;
#pragma unroll 16
  for (int i = 0; i < N; ++i) w = __shfl_up_sync(0xffffffffu, w, 1);
  long long t4 = clock64();
  // independent streams: 8 accumulators (throughput of the pipe for one warp)
  double p[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) p[k] = a + k;
#pragma unroll 4
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k] = __dadd_rn(p[k], b);
  }
  long long t5 = clock64();
  double s = 0; for (int k = 0; k < 8; ++k) s += p[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + y + z + w + s;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = (t3 - t2); cyc[3] = (t4 - t3); cyc[4] = (t5 - t4); cyc[5] = N;
  }
}

__global__ void k_lds(double *out, long long *cyc) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 3) & 1023);
  __syncthreads();
  const int N = 2048;
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) idx = (int)sm[idx & 1023];
  long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = N; }
}

__global__ void k_l2(const long long *chain, long long *out, long long *cyc) {
  const int N = 2048;
  long long idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) {
    long long v; asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(chain + idx) : "memory");
    idx = v;
  }
  long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = N; }
}

// ping-pong: block 0 writes flag[0] = i, block 1 waits for it and writes flag[32] = i, block 0 waits ...  (one lane each)
__global__ void k_pingpong(volatile long long *flag, long long *cyc, int rounds) {
  if (threadIdx.x != 0) return;
  long long t0 = clock64();
  if (blockIdx.x == 0) {
    for (int i = 1; i <= rounds; ++i) {
      asm volatile("st.relaxed.gpu.global.s64 [%0], %1;" ::"l"(flag), "l"((long long)i) : "memory");
      long long v;
      do { asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(flag + 32) : "memory"); } while (v < i);
    }
    cyc[0] = clock64() - t0; cyc[1] = rounds;
  } else if (blockIdx.x == gridDim.x - 1) {
    for (int i = 1; i <= rounds; ++i) {
      long long v;
      do { asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory"); } while (v < i);
      asm volatile("st.relaxed.gpu.global.s64 [%0], %1;" ::"l"(flag + 32), "l"((long long)i) : "memory");
    }
  }
}

int main() {
  double *out; long long *cyc, h[8];
  CK(cudaMalloc(&out, 1 << 20)); CK(cudaMalloc(&cyc, 64));
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SM clock (attr) %d kHz\n", clk);
  for (int warps : {1, 4, 8}) {
    k_chain<<<1, 32 * warps>>>(out, cyc, 1.0, 1e-9, warps);
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 48, cudaMemcpyDeviceToHost));
    printf("warps/SM %d: DADD dep %.1f cyc, DMUL dep %.1f, DMUL+DSUB dep %.1f, SHFL dep %.1f, DADD 8-way indep %.2f cyc/instr\n", warps,
           (double)h[0] / h[5], (double)h[1] / h[5], (double)h[2] / h[5], (double)h[3] / h[5], (double)h[4] / h[5]);
  }
  k_lds<<<1, 32>>>(out, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
  printf("LDS dependent (incl. cvt) %.1f cyc\n", (double)h[0] / h[1]);
  {
    const int M = 1 << 16; long long *hc = (long long *)malloc(M * 8), *dc;
    for (int i = 0; i < M; ++i) hc[i] = (i * 4099LL + 64) % M;
    CK(cudaMalloc(&dc, M * 8)); CK(cudaMemcpy(dc, hc, M * 8, cudaMemcpyHostToDevice));
    k_l2<<<1, 32>>>(dc, (long long *)out, cyc); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
    printf("ld.relaxed.gpu dependent (L2 hit) %.1f cyc\n", (double)h[0] / h[1]);
  }
  {
    long long *flag; CK(cudaMalloc(&flag, 1024));
    for (int nb : {2, 16, 148}) {
      CK(cudaMemset(flag, 0, 1024));
      k_pingpong<<<nb, 32>>>(flag, cyc, 2000); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
      printf("ping-pong block 0 <-> block %d: %.1f cyc per round trip (= 2 hand-offs)\n", nb - 1, (double)h[0] / h[1]);
    }
  }
  return 0;
}

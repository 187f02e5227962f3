// Kernel laboratory for the wave-tile triangular solve: the library's kernel (csrc/wave.cu, included as is)
// on a synthetic 27-point ILU(0)-like factor, checked bit for bit against the sequential CRS_LUSolve loops (CRSMatrix.F90:4642-4660)
// on the host, timed with CUDA events, and traced per task (start / end time stamps, slow-path entries, poll rounds).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -o skew_lab skew_lab.cu
//   ./skew_lab NR NL NP [reps] [trace-file]
// Not part of the product; the product path is lu_apply_wave() in csrc/wave.cu.
#define WAVE_LAB 1
#include "../../elmerfem_b200/csrc/wave.cu"
#include <chrono>
#include <random>

using namespace b200;
namespace b200 { void set_last_error(const std::string &) {} }

static void build_stencil_crs(int NR, int NL, int NP, std::vector<int> &rows, std::vector<int> &cols, std::vector<int> &diag,
                              std::vector<double> &ilu) {
  const long long n = (long long)NR * NL * NP;
  rows.assign(n + 1, 0);
  for (long long i = 0; i < n; ++i) {
    const int a = i % NR, b = (i / NR) % NL, c = i / ((long long)NR * NL);
    int cnt = 0;
    for (int dc = -1; dc <= 1; ++dc) for (int db = -1; db <= 1; ++db) for (int da = -1; da <= 1; ++da)
      if (a + da >= 0 && a + da < NR && b + db >= 0 && b + db < NL && c + dc >= 0 && c + dc < NP) ++cnt;
    rows[i + 1] = rows[i] + cnt;
  }
  cols.resize(rows[n]); ilu.resize(rows[n]); diag.resize(n);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < n; ++i) {
    const int a = i % NR, b = (i / NR) % NL, c = i / ((long long)NR * NL);
    std::mt19937_64 rng(0x9E3779B97F4A7C15ULL ^ (unsigned long long)i);
    std::uniform_real_distribution<double> U(-0.07, 0.07), D(0.5, 1.5);
    int p = rows[i];
    for (int dc = -1; dc <= 1; ++dc) for (int db = -1; db <= 1; ++db) for (int da = -1; da <= 1; ++da) {
      const int a2 = a + da, b2 = b + db, c2 = c + dc;
      if (a2 < 0 || a2 >= NR || b2 < 0 || b2 >= NL || c2 < 0 || c2 >= NP) continue;
      const long long j = a2 + (long long)NR * (b2 + (long long)NL * c2);
      cols[p] = (int)j;
      if (j == i) { diag[i] = p; ilu[p] = D(rng); } else ilu[p] = U(rng);
      ++p;
    }
  }
}

static inline double h_nfms(double a, double b, double c) { volatile double p = b * c; return a - p; }

int main(int argc, char **argv) {
  setvbuf(stdout, nullptr, _IONBF, 0);
  const int NR = argc > 1 ? atoi(argv[1]) : 201, NL = argc > 2 ? atoi(argv[2]) : NR, NP = argc > 3 ? atoi(argv[3]) : NL;
  const int reps = argc > 4 ? atoi(argv[4]) : 5;
  const char *trace_file = argc > 5 ? argv[5] : nullptr;
  const long long n = (long long)NR * NL * NP;
  std::vector<int> rows, cols, diag; std::vector<double> ilu;
  build_stencil_crs(NR, NL, NP, rows, cols, diag, ilu);
  printf("grid %d x %d x %d  n %lld nnz %d\n", NR, NL, NP, n, rows[n]);
  Handle h;
  B200_CUDA(cudaStreamCreate(&h.stream));
  h.n = (int)n; h.nnz = rows[n];
  h.h_rows = rows; h.h_cols = cols; h.h_diag = diag;
  h.d_rows.ensure(n + 1); h.d_cols.ensure(rows[n]); h.d_ilu.ensure(rows[n]); h.ctrl.ensure(1);
  B200_CUDA(cudaMemcpy(h.d_rows.p, rows.data(), (n + 1) * 4, cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(h.d_cols.p, cols.data(), (size_t)rows[n] * 4, cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemcpy(h.d_ilu.p, ilu.data(), (size_t)rows[n] * 8, cudaMemcpyHostToDevice));
  B200_CUDA(cudaMemset(h.ctrl.p, 0, sizeof(Ctrl)));
  h.wv_blocks_per_sm = getenv("B200_WAVE_BLOCKS_PER_SM") ? atoi(getenv("B200_WAVE_BLOCKS_PER_SM")) : 0;
  h.wv_cfg = getenv("B200_WAVE_CFG") ? atoi(getenv("B200_WAVE_CFG")) : 0;
  h.wv_e = getenv("B200_WAVE_E") ? atoi(getenv("B200_WAVE_E")) : 3;
  wave_analyse(h);
  if (!h.wv.ready) { printf("structure not detected\n"); return 2; }
  wave_refresh_values(h);
  B200_CUDA(cudaStreamSynchronize(h.stream));
  const WaveGeom &g = h.wv.g;
  printf("tiles %d x %d lines, %d strips x %d groups, %d tiles of %d steps, layout %.2f x rows\n", g.TB, g.TC, g.NS, g.NG, g.ntiles, g.NT, (double)g.vlen() / n);

  std::vector<double> v(n), ref(n), got(n);
  { std::mt19937_64 rng(7); std::normal_distribution<double> N(0, 1); for (auto &e : v) e = N(rng); }
  // host reference: CRS_LUSolve (unit lower, inverse diagonal stored), separate roundings
  {
    auto t0 = std::chrono::steady_clock::now();
    ref = v;
    for (long long i = 0; i < n; ++i) { double s = ref[i]; for (int p = rows[i]; p < diag[i]; ++p) s = h_nfms(s, ilu[p], ref[cols[p]]); ref[i] = s; }
    for (long long i = n - 1; i >= 0; --i) { double s = ref[i]; for (int p = diag[i] + 1; p < rows[i + 1]; ++p) s = h_nfms(s, ilu[p], ref[cols[p]]); ref[i] = ilu[diag[i]] * s; }
    printf("host reference %.2f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
  DBuf<double> dv, du; dv.ensure(n); du.ensure(n);
  B200_CUDA(cudaMemcpy(dv.p, v.data(), n * 8, cudaMemcpyHostToDevice));
  if (trace_file) wave_trace_enable(h, true);
  lu_apply_wave(h, du.p, dv.p);
  B200_CUDA(cudaStreamSynchronize(h.stream));
  B200_CUDA(cudaMemcpy(got.data(), du.p, n * 8, cudaMemcpyDeviceToHost));
  Ctrl hc; B200_CUDA(cudaMemcpy(&hc, h.ctrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost));
  long long bad = 0, first = -1;
  for (long long i = 0; i < n; ++i) if (memcmp(&got[i], &ref[i], 8)) { if (first < 0) first = i; ++bad; }
  printf("bitwise mismatches %lld (first %lld) spin_timeout %d\n", bad, first, hc.spin_timeout);
  if (trace_file) {
    std::vector<long long> tr; wave_trace_fetch(h, tr);
    FILE *f = fopen(trace_file, "w");
    fprintf(f, "# sweep tile start_ns end_ns polls smid loader0: cyc_mbar_wait cyc_bar cyc_issue -\n");
    const long long nt = g.ntiles;
    long long t0 = -1;
    for (size_t k = 0; k < tr.size() / 8; ++k) if (tr[k * 8 + 0] && (t0 < 0 || tr[k * 8 + 0] < t0)) t0 = tr[k * 8 + 0];
    for (int sw = 0; sw < 2; ++sw) for (long long k = 0; k < nt; ++k) {
      const long long *r = &tr[(sw * nt + k) * 8];
      fprintf(f, "%d %lld %lld %lld %lld %lld %lld %lld %lld %lld\n", sw, k, r[0] - t0, r[1] - t0, r[2], r[3], r[4], r[5], r[6], r[7]);
    }
    fclose(f);
    wave_trace_enable(h, false);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) lu_apply_wave(h, du.p, dv.p);
  cudaEventRecord(e0, h.stream);
  for (int r = 0; r < reps; ++r) lu_apply_wave(h, du.p, dv.p);
  cudaEventRecord(e1, h.stream);
  B200_CUDA(cudaStreamSynchronize(h.stream));
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double per = ms / reps;
  const double bytes = 12.0 * rows[n] + 36.0 * n + 4;
  printf("lu_apply_wave %.3f ms per application  (%.0f GB/s CRS-equivalent, %.3f of 6543.7)\n", per, bytes / per * 1e-6, bytes / per * 1e-6 / 6543.7);
  return bad ? 1 : 0;
}

// makespan (in step units) of the task pipeline on an N^3 27-point grid, natural ordering
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
int main(int argc, char **argv) {
  int N = atoi(argv[1]); int R = atoi(argv[2]); double LAG = atof(argv[3]); int NW = argc > 4 ? atoi(argv[4]) : 1 << 30;
  long long n = (long long)N * N * N;
  auto deps = [&](long long i, long long *d) {
    int a = i % N, b = (i / N) % N, c = i / ((long long)N * N); int k = 0;
    for (int dc = -1; dc <= 0; ++dc) for (int db = -1; db <= 1; ++db) for (int da = -1; da <= 1; ++da) {
      int aa = a + da, bb = b + db, cc = c + dc;
      if (aa < 0 || aa >= N || bb < 0 || bb >= N || cc < 0) continue;
      long long j = aa + (long long)N * (bb + (long long)N * cc);
      if (j < i) d[k++] = j;
    }
    return k;
  };
  std::vector<int> lev(n), step(n);
  std::vector<double> fin(n, 0.0);
  long long ntasks = (n + R - 1) / R; long long tot = 0;
  std::vector<double> task_end(ntasks, 0.0);
  double makespan = 0;
  for (long long t = 0; t < ntasks; ++t) {
    long long q0 = t * R, q1 = std::min(n, q0 + R); int nl = 0;
    for (long long i = q0; i < q1; ++i) { long long d[16]; int k = deps(i, d); int l = 0; for (int m = 0; m < k; ++m) if (d[m] >= q0) l = std::max(l, lev[d[m]] + 1); lev[i] = l; nl = std::max(nl, l + 1); }
    std::vector<std::vector<long long>> byl(nl);
    for (long long i = q0; i < q1; ++i) byl[lev[i]].push_back(i);
    double prev = (t >= NW) ? task_end[t - NW] : 0.0;   // warp reuse
    for (int l = 0; l < nl; ++l) for (size_t c0 = 0; c0 < byl[l].size(); c0 += 32) {
      double start = prev;
      for (size_t m = c0; m < std::min(byl[l].size(), c0 + 32); ++m) { long long i = byl[l][m]; long long d[16]; int k = deps(i, d); for (int x = 0; x < k; ++x) if (d[x] < q0) start = std::max(start, fin[d[x]] + LAG); }
      prev = start + 1.0; ++tot;
      for (size_t m = c0; m < std::min(byl[l].size(), c0 + 32); ++m) fin[byl[l][m]] = prev;
    }
    task_end[t] = prev; makespan = std::max(makespan, prev);
  }
  printf("N %d R %d LAG %.1f NW %d: tasks %lld steps %lld makespan %.0f steps (parallelism %.1f)\n", N, R, LAG, NW, ntasks, tot, makespan, tot / makespan);
}

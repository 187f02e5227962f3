// Discrete-event model of the triangular-solve design DESIGN.md section 7 plans for round 2 ("slabs x strips"):
// forward sweep of the ILU0 factor of the 27-point stencil on an N^3 grid in ElmerGrid's natural numbering.
//
//   slab  = P consecutive z-planes, run by one thread-block cluster (NC clusters resident, slab k on cluster k mod NC,
//           a cluster takes its next slab when the previous one is finished);
//   strip = B consecutive x-lines of every plane of the slab, run by one CTA of the cluster (S = ceil(N/B) CTAs);
//   step  = all rows of one LOCAL dependency level of the strip (level = a + 2 b' + 4 c' inside the strip), processed
//           by the CTA's W warps, 32 rows per warp pass.
// Operands produced by the same CTA come from shared memory (no extra latency beyond the step barrier), operands from
// another strip of the same slab through DSMEM (LD cycles after the producing step ends), operands from the previous
// slab through L2 (LG cycles).  A step costs  CSYNC + ceil(rows / (32 W)) * CROW  cycles and at least
// bytes / BWSM cycles of operand streaming (12 B per matrix entry + 16 B per row for b and x).
//
// Latencies measured on this pool's B200 / in B300_MICROARCH.md: L2 hand-off store->visible->polled 1800 cycles
// (profiles/r01_*trace*), DSMEM 215, shared memory 38; clock 1.965 GHz.  CROW is the open number: 2400 cycles per
// warp step in the task kernel of round 1, ~200 if the consumer loop is the planned LDS.LDS.DMUL.DSUB per entry.
//
// usage: cluster_sptrsv_sim N P B W CROW [CSYNC=40 LD=215 LG=1800 BWSM=24 NC=18]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char **argv) {
  if (argc < 6) { fprintf(stderr, "usage: %s N P B W CROW [CSYNC LD LG BWSM NC]\n", argv[0]); return 1; }
  const int N = atoi(argv[1]), P = atoi(argv[2]), B = atoi(argv[3]), W = atoi(argv[4]);
  const double CROW = atof(argv[5]);
  const double CSYNC = argc > 6 ? atof(argv[6]) : 40, LD = argc > 7 ? atof(argv[7]) : 215, LG = argc > 8 ? atof(argv[8]) : 1800;
  const double BWSM = argc > 9 ? atof(argv[9]) : 24;   // bytes per cycle one SM streams (44 GB/s at 1.965 GHz ~ 22-24)
  const int NC = argc > 10 ? atoi(argv[10]) : 18;
  const double GHZ = 1.965;
  const int S = (N + B - 1) / B, K = (N + P - 1) / P;
  const long long n = (long long)N * N * N;
  std::vector<float> fin((size_t)n, 0.f);             // completion time (cycles) of every row
  auto id = [&](int a, int b, int c) { return (size_t)a + (size_t)N * ((size_t)b + (size_t)N * c); };
  std::vector<double> cta_free((size_t)NC * S, 0.0);  // when CTA s of cluster q finished its previous slab
  double makespan = 0, busy = 0, wait_g = 0, wait_d = 0;
  long long steps = 0;
  // Slabs in order; inside a slab the strips interleave in time, so advance all strips level by level:
  // a strip's local level l can start when its own level l-1 is done and the external operands are there.
  for (int k = 0; k < K; ++k) {
    const int c0 = k * P, c1 = std::min(N, c0 + P), q = k % NC;
    std::vector<double> t(S);                           // current time of every strip's CTA
    for (int s = 0; s < S; ++s) t[s] = cta_free[(size_t)q * S + s];
    const int maxl = (N - 1) + 2 * (B - 1) + 4 * (c1 - c0 - 1);
    // Strips depend on each other only "downwards in time" (line b-1 of the same plane is 2 levels earlier, line b+1
    // of the previous plane 2 levels earlier), so processing local levels in lock-step order over all strips is a
    // valid topological order as long as strip s at level l only needs strip s-1 / s+1 rows finalised earlier in this
    // loop; rows of neighbouring strips at the same GLOBAL level are independent.  Use the global level inside the
    // slab as the outer loop for that reason.
    const int gl_max = (N - 1) + 2 * (N - 1) + 4 * (c1 - c0 - 1);
    (void)maxl;
    for (int gl = 0; gl <= gl_max; ++gl) {
      for (int s = 0; s < S; ++s) {
        const int b0 = s * B, b1 = std::min(N, b0 + B);
        // rows of this strip with a + 2 b + 4 (c - c0) == gl
        int rows = 0; double ready = t[s]; double rg = 0, rd = 0;
        for (int c = c0; c < c1; ++c) for (int b = b0; b < b1; ++b) {
          const int a = gl - 2 * b - 4 * (c - c0);
          if (a < 0 || a >= N) continue;
          ++rows;
          for (int dc = -1; dc <= 0; ++dc) for (int db = -1; db <= 1; ++db) for (int da = -1; da <= 1; ++da) {
            const int aa = a + da, bb = b + db, cc = c + dc;
            if (aa < 0 || aa >= N || bb < 0 || bb >= N || cc < 0) continue;
            if (id(aa, bb, cc) >= id(a, b, c)) continue;
            const double f = fin[id(aa, bb, cc)];
            if (cc < c0) { rg = std::max(rg, f + LG); }                       // previous slab: through L2
            else if (bb < b0 || bb >= b1) { rd = std::max(rd, f + LD); }       // other strip of this slab: DSMEM
            else ready = std::max(ready, (double)f);                           // own strip: shared memory
          }
        }
        if (!rows) continue;
        if (rg > ready) { wait_g += rg - ready; ready = rg; }
        if (rd > ready) { wait_d += rd - ready; ready = rd; }
        const double compute = CSYNC + ((rows + 32 * W - 1) / (32 * W)) * CROW;
        const double stream = rows * (13 * 12.0 + 16.0) / BWSM;
        const double dur = std::max(compute, stream);
        const double end = ready + dur;
        busy += dur; ++steps;
        for (int c = c0; c < c1; ++c) for (int b = b0; b < b1; ++b) {
          const int a = gl - 2 * b - 4 * (c - c0);
          if (a >= 0 && a < N) fin[id(a, b, c)] = (float)end;
        }
        t[s] = end;
      }
    }
    for (int s = 0; s < S; ++s) { cta_free[(size_t)q * S + s] = t[s]; makespan = std::max(makespan, t[s]); }
  }
  const double us = makespan / (GHZ * 1e3);
  const double bytes = (double)n * (13 * 12.0 + 16.0);
  printf("N %d P %d B %d (S %d strips, K %d slabs on %d clusters = %d SMs) W %d CROW %.0f CSYNC %.0f LD %.0f LG %.0f BWSM %.0f: "
         "sweep %.0f us (%.0f GB/s of %.2f GB), CTA busy %.0f%%, steps/CTA-slab %.0f, waits L2 %.0f us DSMEM %.0f us (summed over CTAs)\n",
         N, P, B, S, K, NC, NC * S, W, CROW, CSYNC, LD, LG, BWSM, us, bytes / us / 1e3, bytes / 1e9,
         100.0 * busy / (makespan * NC * S), (double)steps / ((double)K * S), wait_g / (GHZ * 1e3), wait_d / (GHZ * 1e3));
  return 0;
}

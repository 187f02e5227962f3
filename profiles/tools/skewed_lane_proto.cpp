// CPU prototype of the "skewed-lane wavefront" triangular solve specified in DESIGN.md section 7 (round-2 kernel).
// Purpose: pin down the schedule and the operand routing BEFORE the CUDA version exists --
//   * lane j of a strip's warp owns line b0+j of one plane and solves row a = t - 2j at step t;
//   * the operand (a-1, b) is the lane's own result of step t-1 (register), (a-1 | a | a+1, b-1) are lane j-1's results of steps
//     t-3 | t-2 | t-1 (shuffle of a 3-deep history), line b0-1 comes from the neighbouring strip (DSMEM ring in the kernel),
//     the nine operands of plane c-1 from the result vector (L2 in the kernel);
//   * the subtractions run in ascending column order, exactly like CRS_LUSolve's forward loop (CRSMatrix.F90:4642-4649), so the
//     result must be BIT-identical to the sequential solve.
// The emulation executes strips and planes in pipeline order with explicit per-value "produced at global time" stamps and checks
// that no operand is consumed before it exists under the schedule (strip s runs DELTA = 2B+1 steps behind strip s-1, plane c
// 4 steps + HOP behind plane c-1); it then compares with the sequential reference bit for bit, for the forward (L) and the backward
// (U, mirrored skew, inverse diagonal applied last) sweep.
//
//   g++ -O2 -o skew skewed_lane_proto.cpp && ./skew 37 23 11 8      (NR NL NP B; any sizes, strips need not divide NL)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

struct Grid { int NR, NL, NP; long long id(int a, int b, int c) const { return a + (long long)NR * (b + (long long)NL * c); } };

// lower-stencil operand offsets in ASCENDING column order: plane c-1 (b-1,b,b+1 x a-1,a,a+1), plane c (b-1 x a-1,a,a+1), (a-1,b)
static const int OFF[13][3] = {{-1,-1,-1},{0,-1,-1},{1,-1,-1},{-1,0,-1},{0,0,-1},{1,0,-1},{-1,1,-1},{0,1,-1},{1,1,-1},
                               {-1,-1,0},{0,-1,0},{1,-1,0},{-1,0,0}};

int main(int argc, char **argv) {
  Grid G{argc > 1 ? atoi(argv[1]) : 37, argc > 2 ? atoi(argv[2]) : 23, argc > 3 ? atoi(argv[3]) : 11};
  const int B = argc > 4 ? atoi(argv[4]) : 8;
  const long long n = (long long)G.NR * G.NL * G.NP;
  std::mt19937_64 rng(314159265);
  std::uniform_real_distribution<double> U(-0.2, 0.2);
  // values: lv[row][e] for the 13 lower operands (0 where the neighbour does not exist: the pad entries), uv likewise for the
  // mirrored upper stencil, dinv per row
  std::vector<double> lv((size_t)n * 13), uv((size_t)n * 13), dinv((size_t)n), rhs((size_t)n);
  for (auto &v : lv) v = U(rng);
  for (auto &v : uv) v = U(rng);
  for (auto &v : dinv) v = 1.0 + U(rng);
  for (auto &v : rhs) v = U(rng) * 10;
  auto inside = [&](int a, int b, int c) { return a >= 0 && a < G.NR && b >= 0 && b < G.NL && c >= 0 && c < G.NP; };

  for (int sweep = 0; sweep < 2; ++sweep) {
    const bool upper = sweep == 1;
    const std::vector<double> &val = upper ? uv : lv;
    // mirrored coordinates for the backward sweep: A = NR-1-a etc., so that "earlier" always means smaller (A,B,C)
    auto nat = [&](int A, int Bq, int C) { return upper ? G.id(G.NR - 1 - A, G.NL - 1 - Bq, G.NP - 1 - C) : G.id(A, Bq, C); };
    // ---- sequential reference (CRS_LUSolve order): forward ascending columns; backward: columns ascending means the MIRRORED
    // stencil is walked from its far end, i.e. entry order 12..0 in mirrored coordinates
    std::vector<double> ref((size_t)n);
    for (long long q = 0; q < n; ++q) {
      const int A = q % G.NR, Bq = (q / G.NR) % G.NL, C = q / ((long long)G.NR * G.NL);
      const long long i = nat(A, Bq, C);
      double s = rhs[(size_t)i];
      for (int k = 0; k < 13; ++k) {
        const int e = upper ? 12 - k : k;
        const int a2 = A + OFF[e][0], b2 = Bq + OFF[e][1], c2 = C + OFF[e][2];
        if (!inside(a2, b2, c2)) continue;
        s = s - val[(size_t)i * 13 + e] * ref[(size_t)nat(a2, b2, c2)];
      }
      ref[(size_t)i] = upper ? dinv[(size_t)i] * s : s;
    }
    // ---- skewed-lane schedule
    const int S = (G.NL + B - 1) / B, DELTA = 2 * B + 1, HOP = 30;          // HOP: L2 hand-off in units of steps (1800 / 60 cycles)
    const double SENT = std::nan("");
    std::vector<double> out((size_t)n, SENT);
    std::vector<long long> stamp((size_t)n, -1);                             // global step at which a row's result exists
    long long violations = 0, max_step = 0, pad_uses = 0;
    for (int C = 0; C < G.NP; ++C) {
      const long long plane_t0 = (long long)C * (4 + HOP);
      for (int s = 0; s < S; ++s) {
        const int b0 = s * B, nb = std::min(B, G.NL - b0);
        const long long strip_t0 = plane_t0 + (long long)s * DELTA;
        // per-lane registers: own result history h[j][0..2] = results of steps t-1, t-2, t-3
        std::vector<double> h((size_t)nb * 3, 0.0);
        const int nsteps = G.NR + 2 * (nb - 1);
        for (int t = 0; t < nsteps; ++t) {
          std::vector<double> res((size_t)nb, 0.0); std::vector<char> act((size_t)nb, 0);
          const long long now = strip_t0 + t;
          for (int j = 0; j < nb; ++j) {
            const int A = t - 2 * j, Bq = b0 + j;
            if (A < 0 || A >= G.NR) continue;
            act[j] = 1;
            const long long i = nat(A, Bq, C);
            double sacc = rhs[(size_t)i];
            for (int k = 0; k < 13; ++k) {
              const int e = upper ? 12 - k : k;                                 // same operation order as the reference
              const int da = OFF[e][0], db = OFF[e][1], dc = OFF[e][2];
              const int a2 = A + da, b2 = Bq + db, c2 = C + dc;
              double x; bool exists = inside(a2, b2, c2);
              if (!exists) { x = h[(size_t)j * 3]; ++pad_uses; }                // pad entry: value 0 x a finite register
              else if (dc == 0 && db == 0) x = h[(size_t)j * 3];               // (a-1, b): own result of step t-1
              else if (dc == 0 && db == -1 && j > 0) x = h[(size_t)(j - 1) * 3 + (1 - da)];   // lane j-1, age 1 (a+1), 2 (a), 3 (a-1)
              else {                                                             // neighbour strip (DSMEM) or previous plane (L2)
                const long long q2 = nat(a2, b2, c2);
                x = out[(size_t)q2];
                const long long need = stamp[(size_t)q2] + (dc == 0 ? 1 : HOP);  // DSMEM: next step; L2: HOP steps later
                if (stamp[(size_t)q2] < 0 || need > now) ++violations;
              }
              const double v = exists ? val[(size_t)i * 13 + e] : 0.0;
              sacc = sacc - v * x;
              if (exists && dc == 0 && (db == 0 || j > 0)) {                    // register / shuffle operands must be the right rows
                const double want = out[(size_t)nat(a2, b2, c2)];
                if (std::memcmp(&want, &x, 8) != 0) ++violations;
              }
            }
            res[j] = upper ? dinv[(size_t)i] * sacc : sacc;
          }
          for (int j = 0; j < nb; ++j) {                                         // end of step: rotate histories, publish
            h[(size_t)j * 3 + 2] = h[(size_t)j * 3 + 1]; h[(size_t)j * 3 + 1] = h[(size_t)j * 3];
            if (act[j]) {
              h[(size_t)j * 3] = res[j];
              const long long i = nat(t - 2 * j, b0 + j, C);
              out[(size_t)i] = res[j]; stamp[(size_t)i] = now;
            }
            // an inactive lane keeps a finite value in h[.][0] (0.0 initially, its last result afterwards)
          }
          max_step = std::max(max_step, now);
        }
      }
    }
    long long diff = 0;
    for (long long i = 0; i < n; ++i) if (std::memcmp(&out[(size_t)i], &ref[(size_t)i], 8) != 0) ++diff;
    printf("%s sweep on %d x %d x %d, strips of %d lines (%d per plane): %lld rows, bitwise differences %lld, schedule violations %lld, "
           "pad operands %lld, makespan %lld steps (levels %d)\n", upper ? "backward" : "forward", G.NR, G.NL, G.NP, B, S, n, diff,
           violations, pad_uses, max_step + 1, (G.NR - 1) + 2 * (G.NL - 1) + 4 * (G.NP - 1) + 1);
    if (diff || violations) return 1;
  }
  return 0;
}

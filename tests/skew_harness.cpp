// TEST INFRASTRUCTURE.  CPU execution of the skewed-lane triangular solve (elmerfem_b200/csrc/skew.cu, opt-in B200_TRI_MODE=2)
// through the SAME geometry / detection / stream-layout code the CUDA kernel uses (csrc/skewgeom.h): streams are filled with
// sk_fill_row, then every task (plane, strip) is walked step by step, lane by lane, with the kernel's operand routing (own register,
// 3-deep history of lane j-1, result vector for the neighbouring strip and the previous plane).  A value read from the result vector
// before it was written is reported as a hazard (the kernel would spin there; in task order that must never be needed... it is needed
// only from LOWER tasks, which this sequential walk has finished).  The caller compares the result with CRS_LUSolve bit for bit.
//   g++ -O2 -shared -fPIC -o skew_harness.so skew_harness.cpp
#include "../elmerfem_b200/csrc/skewgeom.h"
#include <cmath>
#include <cstring>
#include <vector>
using namespace b200;

static inline double nfms(double a, double b, double c) { volatile double p = b * c; return a - p; }   // separate roundings

// rows/cols/diag 0-based.  geom_out[0..4] = NR, NL, NP, BW, S.  Returns 0 ok, 1 structure not detected, 2 hazard.
extern "C" int skew_emulate(int n, const int *rows, const int *cols, const int *diag, const double *ilu, const double *rhs,
                            double *x_out, int *geom_out) {
  SkewGeom g;
  if (sk_detect(n, rows, cols, diag, g)) return 1;
  geom_out[0] = g.NR; geom_out[1] = g.NL; geom_out[2] = g.NP; geom_out[3] = g.BW; geom_out[4] = g.S;
  std::vector<double> SL((size_t)g.total_steps() * 13 * 32, 0.0), SU((size_t)g.total_steps() * 14 * 32, 0.0);
  for (int i = 0; i < n; ++i) sk_fill_row(g, i, rows, cols, ilu, SL.data(), SU.data());
  const double SENT = std::nan("0x4DEAD");
  std::vector<double> y((size_t)n, SENT), x((size_t)n, SENT);
  int hazards = 0;
  for (int sweep = 0; sweep < 2; ++sweep) {
    const bool UPPER = sweep == 1;
    const int NE = UPPER ? 14 : 13;
    const double *S = UPPER ? SU.data() : SL.data();
    const double *in = UPPER ? y.data() : rhs;
    double *out = UPPER ? x.data() : y.data();
    for (long long k = 0; k < g.ntasks(); ++k) {
      const int C = (int)(k / g.S), s = (int)(k % g.S), B0 = s * g.BW, nb = g.nb(s), nsteps = g.nsteps(s);
      double h0[32] = {0}, h1[32] = {0}, h2[32] = {0};
      for (int t = 0; t < nsteps; ++t) {
        double res[32]; bool actv[32];
        for (int lane = 0; lane < 32; ++lane) {
          const int A = t - 2 * lane, Bq = B0 + lane;
          const bool act = lane < nb && A >= 0 && A < g.NR;
          actv[lane] = act;
          const double *v = S + ((g.step_base(C, s) + t) * NE) * 32 + lane;      // v[e] at v[e * 32]
          double xo[12];
          for (int e = 0; e < 12; ++e) {
            int dA, dB, dC; sk_offset(e, dA, dB, dC);
            const bool ex = act && (e < 9 || lane == 0) && g.inside(A + dA, Bq + dB, C + dC);
            xo[e] = 0.0;
            if (ex) {
              xo[e] = out[g.nat(UPPER, A + dA, Bq + dB, C + dC)];
              if (xo[e] != xo[e]) ++hazards;                                     // still the sentinel
            }
          }
          if (lane > 0) { xo[9] = h2[lane - 1]; xo[10] = h1[lane - 1]; xo[11] = h0[lane - 1]; }
          const long long i = act ? g.nat(UPPER, A, Bq, C) : 0;
          double acc = act ? in[i] : 0.0;
          if (!UPPER) {
            for (int e = 0; e < 12; ++e) acc = nfms(acc, v[e * 32], xo[e]);
            acc = nfms(acc, v[12 * 32], h0[lane]);
          } else {
            acc = nfms(acc, v[12 * 32], h0[lane]);
            for (int e = 11; e >= 0; --e) acc = nfms(acc, v[e * 32], xo[e]);
            acc = v[13 * 32] * acc;
          }
          res[lane] = acc;
        }
        for (int lane = 0; lane < 32; ++lane) {                                   // SIMT: all lanes read before any lane writes
          h2[lane] = h1[lane]; h1[lane] = h0[lane];
          if (actv[lane]) { h0[lane] = res[lane]; out[g.nat(UPPER, t - 2 * lane, B0 + lane, C)] = res[lane]; }
        }
      }
    }
  }
  std::memcpy(x_out, x.data(), (size_t)n * sizeof(double));
  return hazards ? 2 : 0;
}

// ---- concurrency check ---------------------------------------------------------------------------------------------------------------
// NW emulated warps take tasks w, w+NW, ... in increasing order (exactly the kernel's assignment) and are stepped in a pseudo-random
// interleaving; a warp's step completes only when every operand it reads from the result vector is there (the kernel spins), otherwise the
// warp stays where it is.  Returns 0 when all tasks finish and the result equals x_ref bit for bit, 3 on deadlock (a full round in which
// no warp could advance), 4 on a wrong result.
extern "C" int skew_emulate_concurrent(int n, const int *rows, const int *cols, const int *diag, const double *ilu, const double *rhs,
                                       const double *x_ref, int NW, unsigned seed) {
  SkewGeom g;
  if (sk_detect(n, rows, cols, diag, g)) return 1;
  std::vector<double> SL((size_t)g.total_steps() * 13 * 32, 0.0), SU((size_t)g.total_steps() * 14 * 32, 0.0);
  for (int i = 0; i < n; ++i) sk_fill_row(g, i, rows, cols, ilu, SL.data(), SU.data());
  const double SENT = std::nan("0x4DEAD");
  std::vector<double> y((size_t)n, SENT), x((size_t)n, SENT);
  struct Warp { long long k; int t; double h0[32], h1[32], h2[32]; bool done; };
  unsigned long long rng = seed * 2654435761ULL + 88172645463325252ULL;
  auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  for (int sweep = 0; sweep < 2; ++sweep) {
    const bool UPPER = sweep == 1;
    const int NE = UPPER ? 14 : 13;
    const double *S = UPPER ? SU.data() : SL.data();
    const double *in = UPPER ? y.data() : rhs;
    double *out = UPPER ? x.data() : y.data();
    std::vector<Warp> W((size_t)NW);
    for (int w = 0; w < NW; ++w) { W[w].k = w; W[w].t = 0; W[w].done = w >= g.ntasks(); std::memset(W[w].h0, 0, sizeof W[w].h0); std::memset(W[w].h1, 0, sizeof W[w].h1); std::memset(W[w].h2, 0, sizeof W[w].h2); }
    long long remaining = 0; for (int w = 0; w < NW; ++w) remaining += !W[w].done;
    int idle_rounds = 0;
    while (remaining) {
      bool progressed = false;
      const int start = (int)(next() % NW);
      for (int q = 0; q < NW; ++q) {
        Warp &wp = W[(start + q) % NW];
        if (wp.done) continue;
        const int burst = 1 + (int)(next() % 7);                                // a warp runs a few steps, then another one gets the SM
        for (int r = 0; r < burst && !wp.done; ++r) {
          const int C = (int)(wp.k / g.S), s = (int)(wp.k % g.S), B0 = s * g.BW, nb = g.nb(s), nsteps = g.nsteps(s), t = wp.t;
          double res[32]; bool actv[32]; bool ready = true;
          for (int lane = 0; lane < 32 && ready; ++lane) {
            const int A = t - 2 * lane, Bq = B0 + lane;
            const bool act = lane < nb && A >= 0 && A < g.NR;
            actv[lane] = act;
            const double *v = S + ((g.step_base(C, s) + t) * NE) * 32 + lane;
            double xo[12];
            for (int e = 0; e < 12; ++e) {
              int dA, dB, dC; sk_offset(e, dA, dB, dC);
              const bool ex = act && (e < 9 || lane == 0) && g.inside(A + dA, Bq + dB, C + dC);
              xo[e] = 0.0;
              if (ex) { xo[e] = out[g.nat(UPPER, A + dA, Bq + dB, C + dC)]; if (xo[e] != xo[e]) { ready = false; break; } }
            }
            if (!ready) break;
            if (lane > 0) { xo[9] = wp.h2[lane - 1]; xo[10] = wp.h1[lane - 1]; xo[11] = wp.h0[lane - 1]; }
            const long long i = act ? g.nat(UPPER, A, Bq, C) : 0;
            double acc = act ? in[i] : 0.0;
            if (!UPPER) { for (int e = 0; e < 12; ++e) acc = nfms(acc, v[e * 32], xo[e]); acc = nfms(acc, v[12 * 32], wp.h0[lane]); }
            else { acc = nfms(acc, v[12 * 32], wp.h0[lane]); for (int e = 11; e >= 0; --e) acc = nfms(acc, v[e * 32], xo[e]); acc = v[13 * 32] * acc; }
            res[lane] = acc;
          }
          if (!ready) break;                                                     // the kernel would spin here
          for (int lane = 0; lane < 32; ++lane) {
            wp.h2[lane] = wp.h1[lane]; wp.h1[lane] = wp.h0[lane];
            if (actv[lane]) { wp.h0[lane] = res[lane]; out[g.nat(UPPER, t - 2 * lane, B0 + lane, C)] = res[lane]; }
          }
          progressed = true;
          if (++wp.t == nsteps) {
            wp.k += NW; wp.t = 0;
            std::memset(wp.h0, 0, sizeof wp.h0); std::memset(wp.h1, 0, sizeof wp.h1); std::memset(wp.h2, 0, sizeof wp.h2);
            if (wp.k >= g.ntasks()) { wp.done = true; --remaining; }
          }
        }
      }
      if (!progressed) { if (++idle_rounds > 2) return 3; } else idle_rounds = 0;
    }
  }
  for (int i = 0; i < n; ++i) if (std::memcmp(&x[(size_t)i], &x_ref[i], 8) != 0) return 4;
  return 0;
}

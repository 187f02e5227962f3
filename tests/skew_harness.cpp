// TEST INFRASTRUCTURE.  CPU execution of the skewed-lane triangular solve (elmerfem_b200/csrc/skew.cu) through the SAME geometry /
// detection / layout / operand-routing code the CUDA kernel uses (csrc/skewgeom.h): streams are filled with sk_fill_row, vectors are
// permuted into the skewed layout, then every task (plane, strip) is walked step by step, lane by lane, exactly as k_skew does it:
// one fetch per lane and step from the result vector (sk_source), E steps ahead, a history of E + 5 values, the three shuffles
// (near / far / in-plane), the operand windows, the reference's subtraction order.  A fetch that finds the sentinel is where the kernel
// would poll: in sequential task order it is a hazard (an operand from a task that has not run), in the concurrent emulation the warp
// waits.  The caller compares the result with CRS_LUSolve bit for bit.
//   g++ -O2 -ffp-contract=off -shared -fPIC -o skew_harness.so skew_harness.cpp
#include "../elmerfem_b200/csrc/skewgeom.h"
#include <cmath>
#include <cstring>
#include <vector>
using namespace b200;

static inline double nfms(double a, double b, double c) { volatile double p = b * c; return a - p; }   // separate roundings
static const int PAD = 16;                                                                            // steps of padding around the layout

struct Warp {
  long long k = 0; int t = 0; bool done = false, fresh = true;
  int E = 3, HN = 8;
  std::vector<double> H;                      // [32][HN]
  double h0[32], wm[32], w0[32], wp[32], um[32], u0[32], up[32], dm[32], d0[32], dp[32];
  SkewSrc src[32]; SkewTask T;
  void start(const SkewGeom &g, bool upper, long long k_, int E_) {
    k = k_; E = E_; HN = E + 5; T = sk_task(g, upper, k); t = -HN; fresh = false;
    H.assign((size_t)32 * HN, 0.0);
    for (int l = 0; l < 32; ++l) { h0[l] = wm[l] = w0[l] = wp[l] = um[l] = u0[l] = up[l] = dm[l] = d0[l] = dp[l] = 0.0; src[l] = sk_source(g, upper, k, l); }
  }
  double &hist(int lane, int step) { return H[(size_t)lane * HN + ((step % HN) + HN) % HN]; }
};

// One step of one warp; returns false (and changes nothing) when an operand is still the sentinel.  hazards counts those in sequential mode.
static bool warp_step(const SkewGeom &g, bool UPPER, Warp &w, const double *S, const double *RHS, double *Y, bool wait) {
  const int NE = UPPER ? 14 : 13, E = w.E, t = w.t, nb = w.T.nb;
  // the heads needed now must exist (the kernel polls them); the request of this step may still be the sentinel
  double head[32];
  for (int l = 0; l < 32; ++l) {
    const SkewSrc &s = w.src[l];
    const int tau = t + s.off;
    head[l] = 0.0;
    if ((unsigned)(tau - s.lo) < (unsigned)s.len) {
      head[l] = Y[s.idx0 + (long long)tau * s.stride];
      if (head[l] != head[l]) { if (wait) return false; return false; }
    }
  }
  for (int l = 0; l < 32; ++l) {                                   // request of this step (value irrelevant until it becomes the head)
    w.hist(l, t) = 0.0;
    w.hist(l, t - E) = head[l];
  }
  double nearv[32], farv[32], wn[32];
  for (int l = 0; l < 32; ++l) if (l == 30) w.h0[l] = head[l];
  for (int l = 0; l < 32; ++l) {
    const int sn = l < 31 ? l + 1 : 31, sf = l == 0 ? 31 : l - 1, sh = l == 0 ? 30 : l - 1;
    nearv[l] = head[sn];
    farv[l] = w.hist(sf, t - E - 4);
    wn[l] = w.h0[sh];
    if (t - 2 * l + 1 >= g.NR) wn[l] = 0.0;
  }
  for (int l = 0; l < 32; ++l) {
    w.um[l] = w.u0[l]; w.u0[l] = w.up[l]; w.up[l] = farv[l];
    w.dm[l] = w.d0[l]; w.d0[l] = w.dp[l]; w.dp[l] = nearv[l];
    w.wm[l] = w.w0[l]; w.w0[l] = w.wp[l]; w.wp[l] = wn[l];
  }
  if (t >= 0) {
    double res[32]; bool actv[32];
    for (int l = 0; l < 32; ++l) {
      const int A = t - 2 * l;
      const bool act = l < nb && A >= 0 && A < g.NR;
      actv[l] = act; res[l] = 0.0;
      if (!act) continue;
      const long long own = sk_vidx(w.T, UPPER, t, l);              // the row's (step, lane) position, natural addressing
      const double *v = S + (own / 32) * NE * 32 + own % 32;
      const double om = w.hist(l, t - E - 4), o0 = w.hist(l, t - E - 3), op = w.hist(l, t - E - 2);
      double acc = RHS[own];
      const double xo[13] = {w.um[l], w.u0[l], w.up[l], om, o0, op, w.dm[l], w.d0[l], w.dp[l], w.wm[l], w.w0[l], w.wp[l], w.h0[l]};
      if (!UPPER) { for (int e = 0; e < 13; ++e) acc = nfms(acc, v[e * 32], xo[e]); }
      else { for (int e = 12; e >= 0; --e) acc = nfms(acc, v[e * 32], xo[e]); acc = v[13 * 32] * acc; }
      res[l] = acc;
    }
    for (int l = 0; l < 32; ++l) if (actv[l]) { w.h0[l] = res[l]; Y[sk_vidx(w.T, UPPER, t, l)] = res[l]; }
  }
  ++w.t;
  return true;
}

struct Plan {
  SkewGeom g; std::vector<double> SL, SU, yin, y, x;
  double *sl() { return SL.data() + (size_t)PAD * 13 * 32; }
  double *su() { return SU.data() + (size_t)PAD * 14 * 32; }
  double *v(std::vector<double> &a) { return a.data() + (size_t)PAD * 32; }
};
static int make_plan(Plan &P, int n, const int *rows, const int *cols, const int *diag, const double *ilu, const double *rhs) {
  if (sk_detect(n, rows, cols, diag, P.g)) return 1;
  const size_t steps = (size_t)P.g.total_steps() + 2 * PAD;
  P.SL.assign(steps * 13 * 32, 0.0); P.SU.assign(steps * 14 * 32, 0.0);
  for (int i = 0; i < n; ++i) sk_fill_row(P.g, i, rows, cols, ilu, P.sl(), P.su());
  const double SENT = std::nan("0x4DEAD");
  P.yin.assign(steps * 32, 0.0); P.y.assign(steps * 32, 0.0); P.x.assign(steps * 32, 0.0);
  for (long long i = 0; i < P.g.vlen(); ++i) { P.v(P.y)[i] = SENT; P.v(P.x)[i] = SENT; }
  for (int i = 0; i < n; ++i) P.v(P.yin)[P.g.vslot(i)] = rhs[i];
  return 0;
}

// rows/cols/diag 0-based.  geom_out[0..4] = NR, NL, NP, BW, S.  Returns 0 ok, 1 structure not detected, 2 hazard.
extern "C" int skew_emulate(int n, const int *rows, const int *cols, const int *diag, const double *ilu, const double *rhs,
                            double *x_out, int *geom_out, int E) {
  Plan P;
  if (make_plan(P, n, rows, cols, diag, ilu, rhs)) return 1;
  const SkewGeom &g = P.g;
  geom_out[0] = g.NR; geom_out[1] = g.NL; geom_out[2] = g.NP; geom_out[3] = g.BW; geom_out[4] = g.S;
  for (int sweep = 0; sweep < 2; ++sweep) {
    const bool UPPER = sweep == 1;
    const double *S = UPPER ? P.su() : P.sl();
    const double *in = UPPER ? P.v(P.y) : P.v(P.yin);
    double *out = UPPER ? P.v(P.x) : P.v(P.y);
    for (long long k = 0; k < g.ntasks(); ++k) {
      Warp w; w.start(g, UPPER, k, E);
      while (w.t < w.T.nsteps) if (!warp_step(g, UPPER, w, S, in, out, false)) return 2;
    }
  }
  for (int i = 0; i < n; ++i) x_out[i] = P.v(P.x)[g.vslot(i)];
  return 0;
}

// ---- concurrency check ---------------------------------------------------------------------------------------------------------------
// NW emulated warps take tasks w, w+NW, ... in increasing order (exactly the kernel's assignment) and are stepped in a pseudo-random
// interleaving; a warp's step completes only when every operand it fetches from the result vector is there (the kernel polls), otherwise
// the warp stays where it is.  Returns 0 when all tasks finish and the result equals x_ref bit for bit, 3 on deadlock (a full round in
// which no warp could advance), 4 on a wrong result.
extern "C" int skew_emulate_concurrent(int n, const int *rows, const int *cols, const int *diag, const double *ilu, const double *rhs,
                                       const double *x_ref, int NW, unsigned seed, int E) {
  Plan P;
  if (make_plan(P, n, rows, cols, diag, ilu, rhs)) return 1;
  const SkewGeom &g = P.g;
  unsigned long long rng = seed * 2654435761ULL + 88172645463325252ULL;
  auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
  for (int sweep = 0; sweep < 2; ++sweep) {
    const bool UPPER = sweep == 1;
    const double *S = UPPER ? P.su() : P.sl();
    const double *in = UPPER ? P.v(P.y) : P.v(P.yin);
    double *out = UPPER ? P.v(P.x) : P.v(P.y);
    std::vector<Warp> W((size_t)NW);
    long long remaining = 0;
    for (int w = 0; w < NW; ++w) { if (w < g.ntasks()) { W[w].start(g, UPPER, w, E); ++remaining; } else W[w].done = true; }
    int idle_rounds = 0;
    while (remaining) {
      bool progressed = false;
      const int start = (int)(next() % NW);
      for (int q = 0; q < NW; ++q) {
        Warp &wp = W[(start + q) % NW];
        if (wp.done) continue;
        const int burst = 1 + (int)(next() % 7);                                // a warp runs a few steps, then another one gets the SM
        for (int r = 0; r < burst && !wp.done; ++r) {
          if (!warp_step(g, UPPER, wp, S, in, out, true)) break;                  // the kernel would poll here
          progressed = true;
          if (wp.t == wp.T.nsteps) {
            const long long kn = wp.k + NW;
            if (kn >= g.ntasks()) { wp.done = true; --remaining; } else wp.start(g, UPPER, kn, E);
          }
        }
      }
      if (!progressed) { if (++idle_rounds > 2) return 3; } else idle_rounds = 0;
    }
  }
  for (int i = 0; i < n; ++i) { const double xi = P.v(P.x)[g.vslot(i)]; if (std::memcmp(&xi, &x_ref[i], 8) != 0) return 4; }
  return 0;
}

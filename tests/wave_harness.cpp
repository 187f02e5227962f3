// TEST INFRASTRUCTURE.  CPU execution of the wave-tile triangular solve (elmerfem_b200/csrc/wave.cu) through the SAME geometry / layout /
// operand-routing code the CUDA kernel uses (csrc/wavegeom.h; grid detection: csrc/skewgeom.h): the two streams are filled with
// wv_fill_row, the right-hand side is permuted into the tile layout, then every tile is walked step by step, thread by thread, exactly as
// k_wave does it: shared-memory result ring of WV_RING steps, neighbour windows (A: (jb-1, w) one step back, B: (jb, w-1) one step back,
// C: (jb-1, w-1) three back, D: (jb-2, w-1) five back), halo lines fetched from the result vector at the mirrored position, the reference's
// subtraction order.  Tiles run one after the other in processing order; a halo fetch that finds the sentinel is therefore a HAZARD (an
// operand from a tile that has not run -- the tile order would deadlock the kernel).  The caller compares the result with CRS_LUSolve
// bit for bit.
//   g++ -O2 -ffp-contract=off -shared -fPIC -o wave_harness.so wave_harness.cpp
#include "../elmerfem_b200/csrc/wavegeom.h"
#include <cmath>
#include <cstring>
#include <vector>
using namespace b200;

static inline double nfms(double a, double b, double c) { volatile double p = b * c; return a - p; }   // separate roundings
static const unsigned long long SENT = 0x7FF4DEADBEEF0B20ULL;
static inline bool is_sent(double v) { unsigned long long u; memcpy(&u, &v, 8); return u == SENT; }
static inline double sentinel() { double v; memcpy(&v, &SENT, 8); return v; }

struct Thread { double Am = 0, A0 = 0, Bm = 0, B0 = 0, Cm = 0, C0 = 0, Cp = 0, Dm = 0, D0 = 0, Dp = 0, h = 0; };

// one sweep: S / RHS at pos(sweep coordinates), result Q at pos(mirrored sweep coordinates); returns the number of hazards
static long long sweep(const WaveGeom &g, const WaveTiles &T, bool UPPER, const double *S, const double *RHS, double *Q) {
  const int NTHR = g.nthr(), NH = g.nhalo(), NE = UPPER ? 14 : 13, YW = g.TB + 2, YH = g.TC + 1, YSLOT = YW * YH;
  long long hazards = 0;
  std::vector<double> Yr((size_t)WV_RING * YSLOT);
  std::vector<Thread> th(NTHR);
  auto ring = [&](int tau, int w, int jb) -> double & { return Yr[(size_t)(tau & (WV_RING - 1)) * YSLOT + (w + 1) * YW + (jb + 2)]; };
  for (int k = 0; k < g.ntiles; ++k) {
    const int sig = T.sig[k], C = T.grp[k];
    std::fill(Yr.begin(), Yr.end(), 0.0);
    std::fill(th.begin(), th.end(), Thread());
    for (int tau = 0; tau < g.NT; ++tau) {
      std::vector<double> out(NTHR);
      for (int tid = 0; tid < NTHR; ++tid) {
        const int jb = tid % g.TB, w = tid / g.TB;
        const WaveLine ln = wv_line(g, sig, C, jb, w);
        Thread &t = th[tid];
        // windows for this step: C and D were advanced at the end of the previous step in the kernel; here at the start (same values)
        if (tau > 0) {
          const double Cn = ring(tau - 3, w - 1, jb - 1), Dn = ring(tau - 5, w - 1, jb - 2);
          t.Cm = t.C0; t.C0 = t.Cp; t.Cp = Cn; t.Dm = t.D0; t.D0 = t.Dp; t.Dp = Dn;
        }
        const double An = ring(tau - 1, w, jb - 1), Bn = ring(tau - 1, w - 1, jb);
        const double *sp = S + ((long long)k * g.NT + tau) * NE * NTHR + tid;
        const double rv = RHS[((long long)k * g.NT + tau) * NTHR + tid];
        double v[14];
        for (int e = 0; e < NE; ++e) v[e] = sp[(size_t)e * NTHR];
        const int a = tau - ln.tau0;
        const bool active = ln.valid && a >= 0 && a < g.NR;
        double acc;
        if (!UPPER) {
          acc = nfms(rv, v[0], t.Dm); acc = nfms(acc, v[1], t.D0); acc = nfms(acc, v[2], t.Dp);
          acc = nfms(acc, v[3], t.Cm); acc = nfms(acc, v[4], t.C0); acc = nfms(acc, v[5], t.Cp);
          acc = nfms(acc, v[6], t.Bm); acc = nfms(acc, v[7], t.B0); acc = nfms(acc, v[8], Bn);
          acc = nfms(acc, v[9], t.Am); acc = nfms(acc, v[10], t.A0); acc = nfms(acc, v[11], An); acc = nfms(acc, v[12], t.h);
        } else {
          acc = nfms(rv, v[12], t.h); acc = nfms(acc, v[11], An); acc = nfms(acc, v[10], t.A0); acc = nfms(acc, v[9], t.Am); acc = nfms(acc, v[8], Bn);
          acc = nfms(acc, v[7], t.B0); acc = nfms(acc, v[6], t.Bm); acc = nfms(acc, v[5], t.Cp); acc = nfms(acc, v[4], t.C0); acc = nfms(acc, v[3], t.Cm);
          acc = nfms(acc, v[2], t.Dp); acc = nfms(acc, v[1], t.D0); acc = nfms(acc, v[0], t.Dm);
          acc = v[13] * acc;
        }
        if (!active) acc = 0.0;
        out[tid] = acc;
        if (active) Q[wv_pos_mirror(g, T.tile_of.data(), 0, ln.b, ln.c) - (long long)a * NTHR] = acc;
        t.h = acc; t.Am = t.A0; t.A0 = An; t.Bm = t.B0; t.B0 = Bn;
      }
      for (int tid = 0; tid < NTHR; ++tid) ring(tau, tid / g.TB, tid % g.TB) = out[tid];
      for (int hh = 0; hh < NH; ++hh) {
        int jb, w; wv_halo(g, hh, jb, w);
        const WaveLine ln = wv_line(g, sig, C, jb, w);
        const int a = tau - ln.tau0;
        double head = 0.0;
        if (ln.valid && a >= 0 && a < g.NR) {
          head = Q[wv_pos_mirror(g, T.tile_of.data(), 0, ln.b, ln.c) - (long long)a * NTHR];
          if (is_sent(head)) { ++hazards; head = 0.0; }
        }
        ring(tau, w, jb) = head;
      }
    }
  }
  return hazards;
}

extern "C" int wave_emulate(int n, const int *rows, const int *cols, const int *diag, const double *ilu, const double *v, double *x, int *geom_out, int TB,
                            int TC) {
  SkewGeom sg;
  if (sk_detect(n, rows, cols, diag, sg)) return 1;
  WaveGeom g; WaveTiles T;
  wv_plan(g, sg.NR, sg.NL, sg.NP, TB, TC, T);
  geom_out[0] = g.NR; geom_out[1] = g.NL; geom_out[2] = g.NP; geom_out[3] = g.ntiles; geom_out[4] = g.NT;
  // every tile's dependencies have smaller numbers
  const size_t steps = (size_t)g.nsteps(), nthr = (size_t)g.nthr();
  std::vector<double> SL(steps * 13 * nthr, 0.0), SU(steps * 14 * nthr, 0.0), yin(steps * nthr, 0.0), y(steps * nthr, sentinel()), xs(steps * nthr, sentinel());
  for (int i = 0; i < n; ++i) wv_fill_row(g, T.tile_of.data(), i, rows, cols, ilu, SL.data(), SU.data());
  for (int i = 0; i < n; ++i) yin[wv_pos(g, T.tile_of.data(), i % g.NR, (i / g.NR) % g.NL, i / (g.NR * g.NL))] = v[i];
  long long hz = sweep(g, T, false, SL.data(), yin.data(), y.data());
  hz += sweep(g, T, true, SU.data(), y.data(), xs.data());
  for (int i = 0; i < n; ++i) x[i] = xs[wv_pos(g, T.tile_of.data(), i % g.NR, (i / g.NR) % g.NL, i / (g.NR * g.NL))];
  return hz ? 2 : 0;
}

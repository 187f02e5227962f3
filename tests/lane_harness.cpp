// TEST INFRASTRUCTURE.  CPU execution of the lane-tile triangular solve (elmerfem_b200/csrc/lane.cu) through the SAME geometry / layout /
// shuffle-routing / row-arithmetic code the CUDA kernel uses (csrc/lanegeom.h; grid detection: csrc/skewgeom.h): the two streams are
// filled with lt_fill_row, the right-hand side is written into the forward stream's right-hand-side rows, then every tile is walked step by step with 32 lane
// states: lt_send / lt_recv play the two shuffles of a step, lt_row is the row program, replayed values (ghost lanes, plane slot 0) are
// read from the result vector at the mirrored position.  Tiles run one after the other in processing order; a replayed value that is
// still the sentinel is therefore a HAZARD (an operand from a tile that has not run -- the tile order would deadlock the kernel).  The
// caller compares the result with CRS_LUSolve bit for bit.
//   g++ -O2 -ffp-contract=off -shared -fPIC -o lane_harness.so lane_harness.cpp
#include "../elmerfem_b200/csrc/lanegeom.h"
#include <cmath>
#include <cstring>
#include <vector>
using namespace b200;

static const unsigned long long SENT = 0x7FF4DEADBEEF0B20ULL;
static inline bool is_sent(double v) { unsigned long long u; memcpy(&u, &v, 8); return u == SENT; }
static inline double sentinel() { double v; memcpy(&v, &SENT, 8); return v; }

template <bool UPPER, int TC> struct Sweep {
  const LaneGeom &g; const LaneTiles &T; const double *S; double *Q, *R2; long long hazards = 0;   // R2: the other sweep's stream (right-hand-side rows)
  std::vector<LaneHist<TC>> h; std::vector<LaneMsg<TC>> m;
  int k = 0, sig = 0, C = 0;
  Sweep(const LaneGeom &g_, const LaneTiles &T_, const double *S_, double *Q_, double *R2_) : g(g_), T(T_), S(S_), Q(Q_), R2(R2_), h(32), m(32) {}
  template <int U> void step(int tau) {
    constexpr int NE = UPPER ? 14 : 13, NROW = NE + 1;
    for (int j = 0; j < 32; ++j) lt_send<TC, U>(h[j], m[j]);
    for (int j = 31; j >= 0; --j) lt_recv<TC, U>(h[j], m[j ? j - 1 : 0]);
    const long long blk = ((long long)k * g.NT + tau) * TC;
    for (int j = 0; j < 32; ++j) {
      double out[TC + 1];
      for (int p = 0; p <= TC; ++p) {
        const LaneLine ln = lt_line(g, sig, C, j, p);
        const int a = tau - 2 * j - 2 * p;
        const bool active = ln.valid && a >= 0 && a < g.NR;
        double val = 0.0;
        if (lt_replayed(j, p)) {
          if (active) {
            val = Q[lt_pos_mirror(g, T.tile_of.data(), 0, ln.b, ln.c) - (long long)a * g.stride()];
            if (is_sent(val)) { ++hazards; val = 0.0; }
          }
        } else {
          double v[14];
          for (int e = 0; e < NE; ++e) v[e] = S[((blk + p - 1) * NROW + e) * 32 + j];
          const double acc = lt_row<UPPER, TC, U>(h[j], p, v, S[((blk + p - 1) * NROW + NE) * 32 + j]);
          if (active) {
            const long long pm = lt_pos_mirror(g, T.tile_of.data(), 0, ln.b, ln.c) - (long long)a * g.stride();
            val = acc; Q[pm] = acc;
            if (R2) R2[lt_rhs_index(pm, LT_ROWS_U)] = acc;
          }
        }
        out[p] = val;
      }
      for (int p = 0; p <= TC; ++p) h[j].X[p][U & 7] = out[p];
    }
  }
  void run() {
    for (k = 0; k < g.ntiles; ++k) {
      sig = T.sig[k]; C = T.grp[k];
      for (auto &s : h) lt_hist_clear(s);
      for (int t0 = 0; t0 < g.NT; t0 += 8) {
        step<0>(t0); step<1>(t0 + 1); step<2>(t0 + 2); step<3>(t0 + 3); step<4>(t0 + 4); step<5>(t0 + 5); step<6>(t0 + 6); step<7>(t0 + 7);
      }
    }
  }
};

template <int TC> static int emulate(const SkewGeom &sg, int n, const int *rows, const int *cols, const double *ilu, const double *v, double *x, int *geom_out) {
  LaneGeom g; LaneTiles T;
  lt_plan(g, sg.NR, sg.NL, sg.NP, TC, T);
  geom_out[0] = g.NR; geom_out[1] = g.NL; geom_out[2] = g.NP; geom_out[3] = g.ntiles; geom_out[4] = g.NT;
  const size_t nv = (size_t)g.vlen();
  std::vector<double> SL(nv * LT_ROWS_L, 0.0), SU(nv * LT_ROWS_U, 0.0), y(nv, sentinel()), xs(nv, sentinel());
  for (int i = 0; i < n; ++i) lt_fill_row(g, T.tile_of.data(), i, rows, cols, ilu, SL.data(), SU.data());
  for (int i = 0; i < n; ++i) SL[lt_rhs_index(lt_pos(g, T.tile_of.data(), i % g.NR, (i / g.NR) % g.NL, i / (g.NR * g.NL)), LT_ROWS_L)] = v[i];
  Sweep<false, TC> f(g, T, SL.data(), y.data(), SU.data()); f.run();
  Sweep<true, TC> b(g, T, SU.data(), xs.data(), nullptr); b.run();
  for (int i = 0; i < n; ++i) x[i] = xs[lt_pos(g, T.tile_of.data(), i % g.NR, (i / g.NR) % g.NL, i / (g.NR * g.NL))];
  return (f.hazards + b.hazards) ? 2 : 0;
}

extern "C" int lane_emulate(int n, const int *rows, const int *cols, const int *diag, const double *ilu, const double *v, double *x, int *geom_out, int TC) {
  SkewGeom sg;
  if (sk_detect(n, rows, cols, diag, sg)) return 1;
  switch (TC) {
    case 1: return emulate<1>(sg, n, rows, cols, ilu, v, x, geom_out);
    case 2: return emulate<2>(sg, n, rows, cols, ilu, v, x, geom_out);
    case 3: return emulate<3>(sg, n, rows, cols, ilu, v, x, geom_out);
    case 4: return emulate<4>(sg, n, rows, cols, ilu, v, x, geom_out);
  }
  return 3;
}

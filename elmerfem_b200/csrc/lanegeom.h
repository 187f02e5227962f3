// Geometry, data layout, operand routing and row arithmetic of the "lane tile" triangular solve (lane.cu).
// Plain C++ shared by the CUDA kernel, the host planner and the CPU emulation (tests/lane_harness.cpp): the layout, the shuffle routing
// and the row program the kernel runs are the ones the CPU check has executed lane by lane.
//
// The matrix is the ILU(0) factor of the 27-point stencil on an NR x NL x NP grid in natural numbering i = a + NR (b + NL c) (detected and
// verified row by row by sk_detect, skewgeom.h).  Row (a, b, c) of the forward sweep needs (a-1, b, c), the rows a-1 | a | a+1 of line
// (b-1, c) and of the lines (b-1 | b | b+1, c-1); its dependency level is a + 2 b + 4 c.  In the sheared line coordinate beta = b + c the
// level is a + 2 beta + 2 c and the neighbour lines are (beta-1, c) and (beta-2 | beta-1 | beta, c-1).
//
// TILES.  A tile = one WARP = LT_BW = 30 consecutive beta (lanes 2..31; lanes 0, 1 are GHOST lanes that replay the last two lines of
// the strip to the left) x TC consecutive planes held by the same lane.  At tile step tau lane j solves, for every plane slot
// p = 1..TC, row a = tau - 2 j - 2 p of line (beta_j, c_p); plane slot p = 0 is the last plane of the plane group below, replayed.
// All rows of a tile step have the same dependency level, so one step = one level, and every operand of a row is
//   * the lane's own result of the previous step (e12: age 1),
//   * results of lane j-1 (same plane: ages 1, 2, 3 -> e11, e10, e9; plane below: ages 3, 4, 5 -> e5, e4, e3): ONE shuffle per
//     produced value (R = shfl_up(X, 1) one step after it was produced),
//   * results of lane j-2, plane below (ages 5, 6, 7 -> e2, e1, e0): a second shuffle T = shfl_up(R, 1) at age 5,
//   * the lane's own results for the plane below (ages 1, 2, 3 -> e8, e7, e6): registers.
// Nothing inside a tile touches shared memory or a barrier.  Replayed values (ghost lanes, plane slot 0) are read from the result
// vector in global memory a few steps ahead (sentinel protocol: the vector is pre-filled with a NaN payload no arithmetic produces).
// Tile (sigma, C) depends on (sigma-1, C), (sigma, C-1), (sigma-1, C-1): all have a smaller start level 2 BW sigma + 2 TC C; tiles are
// numbered by start level and dealt round-robin to co-resident warps, which cannot deadlock for any grid or warp count.
//
// LAYOUT.  pos(a, b, c) = ((tile * NT + tau) * TC + (p - 1)) * 32 + lane.  A tile step's matrix entries AND right-hand sides are ONE
// contiguous block of the sweep's stream (TC x (13 + 1) rows of 32 doubles forward, TC x (14 + 1) backward: the last row of a plane
// slot is the right-hand side), moved by one bulk copy.  The forward sweep therefore stores each result twice: into the result vector
// (where neighbouring tiles replay it from) and into the right-hand-side row of the backward stream.
//
// SWEEP COORDINATES.  The backward sweep is the same program on the mirrored grid (a, b, c) -> (NR-1-a, NL-1-b, NP-1-c); operand slots
// e = 0..12 as in skewgeom.h (sk_slot): ascending natural column order for the forward sweep, DESCENDING e for the backward sweep.  A
// sweep reads its right-hand side and its matrix stream at pos(sweep coordinates) and writes its result at pos(mirrored sweep
// coordinates), where the other sweep finds it as its right-hand side.
#pragma once
#include "skewgeom.h"
#ifndef __CUDACC_RTC__
#include <algorithm>
#include <vector>
#endif

namespace b200 {

constexpr int LT_BW = 30;      // real lines per tile
constexpr int LT_GH = 2;       // ghost lanes
constexpr int LT_ROWS_L = 14, LT_ROWS_U = 15;   // rows of 32 doubles per (tile step, plane slot) in the forward / backward stream

struct LaneGeom {
  int NR = 0, NL = 0, NP = 0;   // rows per line, lines per plane, planes
  int TC = 0;                   // planes per tile
  int NS = 0, NG = 0;           // strips over beta in [0, NL + NP - 1), plane groups
  int NT = 0;                   // steps per tile (multiple of 8)
  int ntiles = 0;               // non-empty tiles
  SK_HD long long nsteps() const { return (long long)ntiles * NT; }
  SK_HD long long vlen() const { return nsteps() * TC * 32; }
  SK_HD long long stride() const { return (long long)TC * 32; }    // distance of consecutive rows of a line
};

SK_HD inline long long lt_pos(const LaneGeom &g, const int *tile_of, int a, int b, int c) {
  const int beta = b + c, sig = beta / LT_BW, j = beta - sig * LT_BW + LT_GH, C = c / g.TC, p = c - C * g.TC + 1;
  const int k = tile_of[C * g.NS + sig];
  const int tau = a + 2 * j + 2 * p;
  return (((long long)k * g.NT + tau) * g.TC + (p - 1)) * 32 + j;
}
SK_HD inline long long lt_pos_mirror(const LaneGeom &g, const int *tile_of, int a, int b, int c) {
  return lt_pos(g, tile_of, g.NR - 1 - a, g.NL - 1 - b, g.NP - 1 - c);
}
// index of the right-hand-side row entry of position pos in a stream with `rows` rows per plane slot
SK_HD inline long long lt_rhs_index(long long pos, int rows) { return ((pos >> 5) * rows + (rows - 1)) * 32 + (pos & 31); }

// line held by lane j at plane slot p (0 = replayed plane below the group) of tile (sig, C)
struct LaneLine { int b = 0, c = 0; bool valid = false; };
SK_HD inline LaneLine lt_line(const LaneGeom &g, int sig, int C, int j, int p) {
  LaneLine r;
  const int beta = sig * LT_BW + j - LT_GH;
  r.c = C * g.TC + p - 1; r.b = beta - r.c;
  r.valid = r.c >= 0 && r.c < g.NP && r.b >= 0 && r.b < g.NL;
  return r;
}
SK_HD inline bool lt_replayed(int j, int p) { return p == 0 || j < LT_GH; }

// Scatters row i of the ILU factor (CRS order, inverse diagonal stored on the diagonal, CRSMatrix.F90:3654-3660) into the forward stream
// SL (LT_ROWS_L = 14 rows of 32 per plane slot and step: 13 entries + right-hand side) and the backward stream SU (LT_ROWS_U = 15 rows:
// 13 entries, inverse diagonal, right-hand side).
SK_HD inline void lt_fill_row(const LaneGeom &g, const int *tile_of, int i, const int *rows, const int *cols, const double *ilu, double *SL, double *SU) {
  const int a = i % g.NR, b = (i / g.NR) % g.NL, c = i / (g.NR * g.NL);
  const long long pf = lt_pos(g, tile_of, a, b, c), pb = lt_pos_mirror(g, tile_of, a, b, c);
  const long long sf = pf >> 5, sb = pb >> 5;
  const int jf = (int)(pf & 31), jb = (int)(pb & 31);
  for (int q = rows[i]; q < rows[i + 1]; ++q) {
    const int col = cols[q];
    const int da = col % g.NR - a, db = (col / g.NR) % g.NL - b, dc = col / (g.NR * g.NL) - c;
    if (col < i) { const int e = sk_slot(da, db, dc); if (e >= 0) SL[(sf * LT_ROWS_L + e) * 32 + jf] = ilu[q]; }
    else if (col > i) { const int e = sk_slot(-da, -db, -dc); if (e >= 0) SU[(sb * LT_ROWS_U + e) * 32 + jb] = ilu[q]; }
    else SU[(sb * LT_ROWS_U + 13) * 32 + jb] = ilu[q];
  }
}

// ---- the lane program ----------------------------------------------------------------------------------------------------------------
// Histories as rings of 8 indexed by the step (mod 8) that PRODUCED the value; the step loop is unrolled by 8 so that every index is a
// compile-time constant (registers).  Plane slot p = 0..TC.
template <int TC> struct LaneHist { double X[TC + 1][8], R[TC + 1][8], T[TC][8]; };
template <int TC> struct LaneMsg { double r[TC + 1], t[TC]; };
template <int TC> SK_HD inline void lt_hist_clear(LaneHist<TC> &h) {
  for (int p = 0; p <= TC; ++p) for (int q = 0; q < 8; ++q) { h.X[p][q] = 0.0; h.R[p][q] = 0.0; if (p < TC) h.T[p][q] = 0.0; }
}
// what the lane passes to lane j+1 at the start of step U: its results of step U-1, and what it received (from lane j-1) for step U-5
template <int TC, int U> SK_HD inline void lt_send(const LaneHist<TC> &h, LaneMsg<TC> &m) {
#pragma unroll
  for (int p = 0; p <= TC; ++p) m.r[p] = h.X[p][(U + 7) & 7];
#pragma unroll
  for (int p = 0; p < TC; ++p) m.t[p] = h.R[p][(U + 3) & 7];
}
template <int TC, int U> SK_HD inline void lt_recv(LaneHist<TC> &h, const LaneMsg<TC> &m) {
#pragma unroll
  for (int p = 0; p <= TC; ++p) h.R[p][(U + 7) & 7] = m.r[p];
#pragma unroll
  for (int p = 0; p < TC; ++p) h.T[p][(U + 3) & 7] = m.t[p];
}
// separate roundings: the reference build does not contract a - b*c
SK_HD inline double lt_nfms(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __dsub_rn(a, __dmul_rn(b, c));
#else
  volatile double p = b * c; return a - p;
#endif
}
// Row of plane slot p at step U: v[0..12] entries in slot order, v[13] inverse diagonal (backward), rhs; the reference's left-to-right
// order (forward, CRSMatrix.F90:4642-4649: ascending e; backward, 4653-4660: descending e, inverse diagonal last).
template <bool UPPER, int TC, int U> SK_HD inline double lt_row(const LaneHist<TC> &h, int p, const double *v, double rhs) {
  const int q = p - 1;
  double x[13];
  x[0] = h.T[q][(U + 1) & 7]; x[1] = h.T[q][(U + 2) & 7]; x[2] = h.T[q][(U + 3) & 7];       // lane j-2, plane below: ages 7, 6, 5
  x[3] = h.R[q][(U + 3) & 7]; x[4] = h.R[q][(U + 4) & 7]; x[5] = h.R[q][(U + 5) & 7];       // lane j-1, plane below: ages 5, 4, 3
  x[6] = h.X[q][(U + 5) & 7]; x[7] = h.X[q][(U + 6) & 7]; x[8] = h.X[q][(U + 7) & 7];       // own lane, plane below: ages 3, 2, 1
  x[9] = h.R[p][(U + 5) & 7]; x[10] = h.R[p][(U + 6) & 7]; x[11] = h.R[p][(U + 7) & 7];     // lane j-1, same plane: ages 3, 2, 1
  x[12] = h.X[p][(U + 7) & 7];                                                             // own previous row
  double acc = rhs;
  if (!UPPER) {
#pragma unroll
    for (int e = 0; e < 13; ++e) acc = lt_nfms(acc, v[e], x[e]);
  } else {
#pragma unroll
    for (int e = 12; e >= 0; --e) acc = lt_nfms(acc, v[e], x[e]);
#ifdef __CUDA_ARCH__
    acc = __dmul_rn(v[13], acc);
#else
    acc = v[13] * acc;
#endif
  }
  return acc;
}

#ifndef __CUDACC_RTC__
// Host planner: tile tables for a grid.  Tiles are numbered by (start level, C): every dependency of a tile has a smaller number.
struct LaneTiles { std::vector<int> tile_of, sig, grp; };
inline void lt_plan(LaneGeom &g, int NR, int NL, int NP, int TC, LaneTiles &T) {
  g.NR = NR; g.NL = NL; g.NP = NP; g.TC = TC;
  g.NS = (NL + NP - 1 + LT_BW - 1) / LT_BW; g.NG = (NP + TC - 1) / TC;
  g.NT = (NR + 2 * 31 + 2 * TC + 7) / 8 * 8;
  T.tile_of.assign((size_t)g.NS * g.NG, -1); T.sig.clear(); T.grp.clear();
  std::vector<std::pair<long long, std::pair<int, int>>> order;        // (start level, (C, sigma))
  for (int C = 0; C < g.NG; ++C) for (int s = 0; s < g.NS; ++s) {
    bool any = false;
    for (int w = 0; w < TC && !any; ++w) {
      const int c = C * TC + w;
      if (c >= NP) break;
      const int lo = std::max(c, s * LT_BW), hi = std::min(c + NL, (s + 1) * LT_BW);   // beta range of plane c inside the strip
      any = lo < hi;
    }
    if (any) order.push_back({2LL * LT_BW * s + 2LL * TC * C, {C, s}});
  }
  std::sort(order.begin(), order.end());
  for (size_t k = 0; k < order.size(); ++k) {
    const int C = order[k].second.first, s = order[k].second.second;
    T.tile_of[(size_t)C * g.NS + s] = (int)k; T.sig.push_back(s); T.grp.push_back(C);
  }
  g.ntiles = (int)order.size();
}
#endif

}  // namespace b200

// Geometry, data layout and operand routing of the "wave tile" triangular solve (wave.cu).
// Plain C++ shared by the CUDA kernel, the host planner and the CPU emulation (tests/wave_harness.cpp), so that the layout and the
// routing the kernel uses are the ones the CPU check has executed.
//
// The matrix is the ILU(0) factor of the 27-point stencil on an NR x NL x NP grid in natural numbering i = a + NR (b + NL c)
// (detected and verified row by row by sk_detect, skewgeom.h).  Row (a, b, c) of the forward sweep needs (a-1, b, c), the three rows
// a-1 | a | a+1 of line (b-1, c) and of the lines (b-1 | b | b+1, c-1): its dependency level is a + 2 b + 4 c.
//
// LINES AND STEPS.  One thread owns one line (b, c) and walks it in a; in the sheared line coordinate beta = b + c the level is
// a + 2 beta + 2 c and the four neighbour lines are (beta-1, c), (beta-2 | beta-1 | beta, c-1): every dependency points to a smaller
// or equal beta AND a smaller or equal c.
//
// TILES.  A tile = TB consecutive beta x TC consecutive planes = TB * TC threads of one CTA that advance in lockstep, one row per
// thread and step, one CTA barrier per step: thread (jb, w) solves row a = tau - WV_PRE - 2 jb - 2 w at tile step tau and publishes it
// in a shared-memory ring, where its neighbours find it one (lines (jb-1, w) and (jb, w-1)), three ((jb-1, w-1)) or five
// ((jb-2, w-1)) steps later.  Hand-offs inside a tile never touch L2.  The lines just outside the tile (beta = -2, -1 and plane
// w = -1: TB + 2 + 2 TC "halo" lines) belong to the tiles (sigma-1, C), (sigma, C-1), (sigma-1, C-1): loader threads fetch their rows
// from the result vector in global memory a few steps ahead and publish them in the same ring at the step the owner would have.
// Tiles therefore form a DAG in which (sigma, C) depends only on tiles with smaller sigma + C; a co-resident grid takes them in
// that order, which cannot deadlock for any grid shape or CTA count.
//
// LAYOUT.  pos(a, b, c) = ((tile * NT + tau) * NTHR + thread): the NTHR threads of a tile step are contiguous, so matrix entries and
// right-hand sides of a step are ONE contiguous block (moved by one TMA bulk copy each) and result stores are coalesced.  The matrix
// stream holds, per (tile, step), 13 (forward) or 14 (backward: slot 13 = inverse diagonal) rows of NTHR entries.
//
// SWEEP COORDINATES.  The backward sweep is the same program on the mirrored grid (a, b, c) -> (NR-1-a, NL-1-b, NP-1-c), operand
// slots e = 0..12 as in skewgeom.h (sk_slot): ascending natural column order for the forward sweep, DESCENDING e for the backward
// sweep.  A sweep reads its right-hand side and its matrix stream at pos(sweep coordinates) and writes its result at
// pos(mirrored sweep coordinates) -- which is where the other sweep's TMA copies find it as their right-hand side.
#pragma once
#include "skewgeom.h"
#ifndef __CUDACC_RTC__
#include <algorithm>
#include <vector>
#endif

namespace b200 {

constexpr int WV_PRE = 6;    // a tile's halo line (jb, w) = (-2, -1) publishes its row 0 at step WV_PRE - 6 = 0
constexpr int WV_RING = 8;   // depth of the shared-memory result ring (oldest read: 5 steps back)

struct WaveGeom {
  int NR = 0, NL = 0, NP = 0;   // rows per line, lines per plane, planes
  int TB = 0, TC = 0;           // tile: TB sheared lines x TC planes
  int NS = 0, NG = 0;           // strips over beta in [0, NL + NP - 1), plane groups
  int NT = 0;                   // steps per tile
  int ntiles = 0;               // non-empty tiles
  SK_HD int nthr() const { return TB * TC; }
  SK_HD int nhalo() const { return TB + 2 + 2 * TC; }
  SK_HD long long nsteps() const { return (long long)ntiles * NT; }
  SK_HD long long vlen() const { return nsteps() * nthr(); }
};

// position of row (a, b, c) (sweep coordinates) in the tile layout; tile_of[C * NS + sigma] = tile number (processing order) or -1
SK_HD inline long long wv_pos(const WaveGeom &g, const int *tile_of, int a, int b, int c) {
  const int beta = b + c, sig = beta / g.TB, jb = beta - sig * g.TB, C = c / g.TC, w = c - C * g.TC;
  const int k = tile_of[C * g.NS + sig];
  const int tau = WV_PRE + a + 2 * jb + 2 * w;
  return ((long long)k * g.NT + tau) * g.nthr() + w * g.TB + jb;
}
SK_HD inline long long wv_pos_mirror(const WaveGeom &g, const int *tile_of, int a, int b, int c) {
  return wv_pos(g, tile_of, g.NR - 1 - a, g.NL - 1 - b, g.NP - 1 - c);
}

// a line of tile (sig, C): compute threads jb in [0, TB), w in [0, TC); halo lines jb in {-2, -1} or w = -1
struct WaveLine {
  int b = 0, c = 0;
  int tau0 = 0;          // tile step at which the line's row 0 is published
  bool valid = false;    // the line exists in the grid
};
SK_HD inline WaveLine wv_line(const WaveGeom &g, int sig, int C, int jb, int w) {
  WaveLine r;
  const int beta = sig * g.TB + jb;
  r.c = C * g.TC + w; r.b = beta - r.c;
  r.tau0 = WV_PRE + 2 * jb + 2 * w;
  r.valid = r.c >= 0 && r.c < g.NP && r.b >= 0 && r.b < g.NL;
  return r;
}
// halo line number hh in [0, nhalo) -> (jb, w)
SK_HD inline void wv_halo(const WaveGeom &g, int hh, int &jb, int &w) {
  if (hh < g.TB + 2) { jb = hh - 2; w = -1; }
  else { const int q = hh - (g.TB + 2); jb = -2 + (q & 1); w = q >> 1; }
}

// Scatters row i of the ILU factor (CRS order, inverse diagonal stored on the diagonal, CRSMatrix.F90:3654-3660) into the forward
// stream SL (13 rows of NTHR per step) and the backward stream SU (14 rows, row 13 = inverse diagonal).
SK_HD inline void wv_fill_row(const WaveGeom &g, const int *tile_of, int i, const int *rows, const int *cols, const double *ilu, double *SL, double *SU) {
  const int a = i % g.NR, b = (i / g.NR) % g.NL, c = i / (g.NR * g.NL);
  const int nthr = g.nthr();
  const long long pf = wv_pos(g, tile_of, a, b, c), pb = wv_pos_mirror(g, tile_of, a, b, c);
  const long long sf = pf / nthr, sb = pb / nthr;
  const int tf = (int)(pf - sf * nthr), tb = (int)(pb - sb * nthr);
  for (int q = rows[i]; q < rows[i + 1]; ++q) {
    const int col = cols[q];
    const int da = col % g.NR - a, db = (col / g.NR) % g.NL - b, dc = col / (g.NR * g.NL) - c;
    if (col < i) { const int e = sk_slot(da, db, dc); if (e >= 0) SL[(sf * 13 + e) * nthr + tf] = ilu[q]; }
    else if (col > i) { const int e = sk_slot(-da, -db, -dc); if (e >= 0) SU[(sb * 14 + e) * nthr + tb] = ilu[q]; }
    else SU[(sb * 14 + 13) * nthr + tb] = ilu[q];
  }
}

#ifndef __CUDACC_RTC__
// Host planner: tile tables for a grid.  Tiles are numbered by (sigma + C, C): every dependency of a tile has a smaller number.
struct WaveTiles { std::vector<int> tile_of, sig, grp; };
inline void wv_plan(WaveGeom &g, int NR, int NL, int NP, int TB, int TC, WaveTiles &T) {
  g.NR = NR; g.NL = NL; g.NP = NP; g.TB = TB; g.TC = TC;
  g.NS = (NL + NP - 1 + TB - 1) / TB; g.NG = (NP + TC - 1) / TC;
  g.NT = WV_PRE + NR + 2 * (TB - 1) + 2 * (TC - 1);
  T.tile_of.assign((size_t)g.NS * g.NG, -1); T.sig.clear(); T.grp.clear();
  std::vector<std::pair<int, int>> order;                      // (sigma + C, C)
  for (int C = 0; C < g.NG; ++C) for (int s = 0; s < g.NS; ++s) {
    bool any = false;
    for (int w = 0; w < TC && !any; ++w) {
      const int c = C * TC + w;
      if (c >= NP) break;
      const int lo = std::max(c, s * TB), hi = std::min(c + NL, (s + 1) * TB);   // beta range of plane c inside the strip
      any = lo < hi;
    }
    if (any) order.push_back({s + C, C});
  }
  std::sort(order.begin(), order.end());
  for (size_t k = 0; k < order.size(); ++k) {
    const int C = order[k].second, s = order[k].first - C;
    T.tile_of[(size_t)C * g.NS + s] = (int)k; T.sig.push_back(s); T.grp.push_back(C);
  }
  g.ntiles = (int)order.size();
}
#endif

}  // namespace b200

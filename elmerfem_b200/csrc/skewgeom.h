// Geometry, data layout and operand routing of the "skewed-lane wavefront" triangular solve.
// Plain C++ shared by the CUDA kernel (skew.cu), the host planner and the CPU emulation (tests/skew_harness.cpp), so that the layout
// and the routing the kernel uses are the ones the CPU check has executed.
//
// The matrix must be the ILU(0) factor of a 27-point (or smaller: 9-point, 3-point) stencil on an NR x NL x NP grid in natural
// numbering, i = a + NR (b + NL c) -- the numbering ElmerGrid gives structured meshes.
//
// TASKS.  A plane c is cut into S strips of <= 29 consecutive lines; a task = (plane, strip), run by one warp.  Lane j owns line
// B0 + j of the strip and is skewed by two steps per lane: at step t it solves row a = t - 2 j of its line.  With that skew every
// in-plane operand is in the warp when it is needed: (a-1, b) is the lane's own previous result, (a-1 | a | a+1, b-1) are lane j-1's
// results of the last three steps.
//
// SKEWED LAYOUT.  Row (a, b, c) of a solve vector lives at  slot = (step_base(c, s) + a + 2 j) * 32 + j  (s = b / BW, j = b - s BW):
// the 32 lanes of a task step are contiguous, so every access of the sweep is a coalesced 256-byte row, and the operands of the
// previous plane are the SAME lanes (+-1) of the task (c-1, s) at steps t+3, t+1, t-1.  The matrix entries use the same (task, step,
// lane) position: stream index = ((step_base + t) * NE + e) * 32 + j, NE = 13 (L) or 14 (U, slot 13 = inverse diagonal).
//
// SWEEP COORDINATES.  The forward sweep walks tasks, steps and lanes upwards.  The backward sweep is the same program in mirrored
// coordinates: sweep plane C = NP-1-c, sweep strip s' = S-1-s, sweep step t' = nsteps-1-t, sweep lane j' = nb-1-j, so that a row's
// operands always have smaller sweep coordinates.  Operand slots e = 0..12 are the 13 earlier neighbours in sweep coordinates,
//   e = (dA+1) + 3 (dB+1) for dC = -1;  e = 10 + dA for dC = 0, dB = -1;  e = 12 for (dA, dB, dC) = (-1, 0, 0):
// ascending natural column order for the forward sweep; for the backward sweep DESCENDING e is ascending column order.
#pragma once
#ifndef __CUDACC__
#define SK_HD
#else
#define SK_HD __host__ __device__
#endif

namespace b200 {

constexpr int SK_MAX_LINES = 29;   // lanes 29..31 of a warp fetch the operands that live in neighbouring strips

struct SkewGeom {
  int NR = 0, NL = 0, NP = 0;     // rows per line, lines per plane, planes
  int BW = 0, S = 0;              // lines per strip (<= 29), strips per plane
  int steps_full = 0;             // steps of a full strip: NR + 2 (BW - 1)
  int steps_plane = 0;            // sum over the strips of a plane
  SK_HD int nb(int s) const { int r = NL - s * BW; return r < BW ? r : BW; }
  SK_HD int nsteps(int s) const { return NR + 2 * (nb(s) - 1); }
  SK_HD long long step_base(int c, int s) const { return (long long)c * steps_plane + (long long)s * steps_full; }
  SK_HD long long ntasks() const { return (long long)NP * S; }
  SK_HD long long total_steps() const { return (long long)NP * steps_plane; }
  SK_HD long long vlen() const { return total_steps() * 32; }                                  // doubles of a skewed vector
  SK_HD long long vslot(int a, int b, int c) const { const int s = b / BW, j = b - s * BW; return (step_base(c, s) + a + 2 * j) * 32 + j; }
  SK_HD long long vslot(long long i) const { return vslot((int)(i % NR), (int)((i / NR) % NL), (int)(i / ((long long)NR * NL))); }
};

// slot of an offset in sweep coordinates, -1 if the offset is not one of the 13 earlier neighbours
SK_HD inline int sk_slot(int dA, int dB, int dC) {
  if (dA < -1 || dA > 1 || dB < -1 || dB > 1) return -1;
  if (dC == -1) return (dA + 1) + 3 * (dB + 1);
  if (dC != 0) return -1;
  if (dB == -1) return 10 + dA;
  if (dB == 0 && dA == -1) return 12;
  return -1;
}

// Scatters row i of the ILU factor (CRS order, inverse diagonal stored on the diagonal, CRSMatrix.F90:3654-3660) into the forward
// stream SL (13 slots per step) and the backward stream SU (14 slots, slot 13 = inverse diagonal), both at the row's (task, step, lane).
SK_HD inline void sk_fill_row(const SkewGeom &g, int i, const int *rows, const int *cols, const double *ilu, double *SL, double *SU) {
  const int a = i % g.NR, b = (i / g.NR) % g.NL, c = i / (g.NR * g.NL);
  const int s = b / g.BW, j = b - s * g.BW;
  const long long p = g.step_base(c, s) + a + 2 * j;
  for (int q = rows[i]; q < rows[i + 1]; ++q) {
    const int col = cols[q];
    const int da = col % g.NR - a, db = (col / g.NR) % g.NL - b, dc = col / (g.NR * g.NL) - c;
    if (col < i) { const int e = sk_slot(da, db, dc); if (e >= 0) SL[(p * 13 + e) * 32 + j] = ilu[q]; }
    else if (col > i) { const int e = sk_slot(-da, -db, -dc); if (e >= 0) SU[(p * 14 + e) * 32 + j] = ilu[q]; }
    else SU[(p * 14 + 13) * 32 + j] = ilu[q];
  }
}

// ---- tasks and operand sources in sweep coordinates ---------------------------------------------------------------------------------
struct SkewTask {
  int C = 0, s = 0;               // sweep plane, sweep strip
  int nb = 0, nsteps = 0;
  long long base = 0;             // first step of the task in the layout (natural addressing)
};
SK_HD inline SkewTask sk_task(const SkewGeom &g, bool upper, long long k) {
  SkewTask T;
  T.C = (int)(k / g.S); T.s = (int)(k % g.S);
  const int c = upper ? g.NP - 1 - T.C : T.C, sn = upper ? g.S - 1 - T.s : T.s;
  T.nb = g.nb(sn); T.nsteps = g.nsteps(sn); T.base = g.step_base(c, sn);
  return T;
}
// layout index of (sweep step tau, sweep lane lam) of a task, and the index step per sweep step (+-32)
SK_HD inline long long sk_vidx(const SkewTask &T, bool upper, int tau, int lam) {
  return upper ? (T.base + T.nsteps - 1 - tau) * 32 + (T.nb - 1 - lam) : (T.base + tau) * 32 + lam;
}
// matrix stream: the block of G steps starting at sweep step t0 is the contiguous run of steps [first, first + G) in natural addressing
SK_HD inline long long sk_group_first(const SkewTask &T, bool upper, int t0, int G) { return upper ? T.base + T.nsteps - t0 - G : T.base + t0; }

// What a lane fetches from the result vector, one value per step: the value needed E' steps later sits at
//   index = idx0 + tau * stride,  tau = t + off,  and exists iff 0 <= tau - lo < len.
//   lanes 0..nb-1: their own line in the previous plane, row a + 3 (operand (a+1, b+1) of lane j-1 now, own (a+1, b) two steps later,
//                  (a+1, b-1) of lane j+1 four steps later);
//   lane nb      : first line of the next strip in the previous plane (feeds lane nb-1 like a lane nb would);
//   lane 31      : last line of the previous strip in the previous plane (feeds lane 0 like a lane -1 would, i.e. four steps later);
//   lane 30      : last line of the previous strip in THIS plane (its value is lane 0's in-plane neighbour (a+1, b-1) of this step).
struct SkewSrc { long long idx0 = 0; int stride = 0, off = 0, lo = 0, len = 0; };
SK_HD inline SkewSrc sk_source(const SkewGeom &g, bool upper, long long k, int lane) {
  SkewSrc r;
  const SkewTask T = sk_task(g, upper, k);
  long long ks = -1; int lam = 0;
  if (lane < T.nb) { if (T.C >= 1) { ks = k - g.S; lam = lane; r.off = 3; } }
  else if (lane == T.nb) { if (T.C >= 1 && T.s + 1 < g.S) { ks = k - g.S + 1; lam = 0; r.off = 3 - 2 * T.nb; } }
  else if (lane == 31) { if (T.C >= 1 && T.s >= 1) { ks = k - g.S - 1; const SkewTask P = sk_task(g, upper, ks); lam = P.nb - 1; r.off = 2 * P.nb + 3; } }
  else if (lane == 30) { if (T.s >= 1) { ks = k - 1; const SkewTask P = sk_task(g, upper, ks); lam = P.nb - 1; r.off = 2 * P.nb - 1; } }
  if (ks < 0) return r;
  const SkewTask P = sk_task(g, upper, ks);
  r.idx0 = sk_vidx(P, upper, 0, lam);
  r.stride = upper ? -32 : 32;
  r.lo = 2 * lam; r.len = g.NR;
  return r;
}

inline void sk_set_strips(SkewGeom &g, int bw_max = SK_MAX_LINES) {
  int S = (g.NL + bw_max - 1) / bw_max;
  if (S < 1) S = 1;
  g.BW = (g.NL + S - 1) / S;                 // equal strips: 201 lines -> 7 strips of 29 (last 27)
  if (g.BW < 1) g.BW = 1;
  g.S = (g.NL + g.BW - 1) / g.BW;
  g.steps_full = g.NR + 2 * (g.BW - 1);
  g.steps_plane = 0;
  for (int s = 0; s < g.S; ++s) g.steps_plane += g.nsteps(s);
}

// Detects the grid from a 0-based CRS pattern with sorted columns and verifies that EVERY row's strictly lower and strictly upper
// pattern is exactly the set of existing stencil neighbours.  Returns nullptr on success, else the reason.
inline const char *sk_detect(int n, const int *rows, const int *cols, const int *diag, SkewGeom &g) {
  if (n <= 0) return "empty matrix";
  auto has = [&](int i, int c) { for (int p = rows[i]; p < rows[i + 1]; ++p) if (cols[p] == c) return true; return false; };
  int NR = n;
  for (int i = 1; i < n; ++i) if (!has(i, i - 1)) { NR = i; break; }
  if (n % NR) return "row count is not a multiple of the line length";
  const int nlines = n / NR;
  int NL = nlines;
  for (int b = 1; b < nlines; ++b) if (!has(b * NR, (b - 1) * NR)) { NL = b; break; }
  if (nlines % NL) return "line count is not a multiple of the plane size";
  g.NR = NR; g.NL = NL; g.NP = nlines / NL;
  bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok)
  for (int i = 0; i < n; ++i) {
    const int a = i % NR, b = (i / NR) % NL, c = i / (NR * NL);
    if (cols[diag[i]] != i) { ok = false; continue; }
    int p = rows[i];
    for (int dc = -1; dc <= 1 && ok; ++dc) for (int db = -1; db <= 1; ++db) for (int da = -1; da <= 1; ++da) {
      const int a2 = a + da, b2 = b + db, c2 = c + dc;
      if (a2 < 0 || a2 >= NR || b2 < 0 || b2 >= NL || c2 < 0 || c2 >= g.NP) continue;
      const long long j = a2 + (long long)NR * (b2 + (long long)NL * c2);
      if (p >= rows[i + 1] || cols[p] != j) { ok = false; break; }
      ++p;
    }
    if (p != rows[i + 1]) ok = false;
  }
  if (!ok) return "a row's pattern is not the full 27-point stencil of the detected grid";
  sk_set_strips(g);
  return nullptr;
}

}  // namespace b200

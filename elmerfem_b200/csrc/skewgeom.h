// Geometry and stream layout of the "skewed-lane wavefront" triangular solve (opt-in, B200_TRI_MODE=2).
// Plain C++ shared by the CUDA kernel (skew.cu), the host planner and the CPU test harness (tests/skew_harness.cpp), so that the
// layout the kernel reads is the layout the CPU check has executed.
//
// The matrix must be the ILU(0) factor of a 27-point (or smaller: 9-point, 3-point) stencil on an NR x NL x NP grid in natural
// numbering, i = a + NR (b + NL c) -- the numbering ElmerGrid gives structured meshes.  A sweep walks "sweep coordinates" (A, B, C):
// the natural ones for the forward sweep, the mirrored ones (NR-1-a, ...) for the backward sweep, so that a row's operands always
// have smaller coordinates.  Operand slots e = 0..12 are the 13 earlier neighbours in ASCENDING natural column order for the
// forward sweep (and, mirrored, DESCENDING order of e is ascending column order for the backward sweep):
//   e = (da+1) + 3 (db+1) + 9 (dc+1) restricted to the 13 offsets below; slot 13 of the backward stream holds the inverse diagonal.
// One warp runs a strip of BW <= 32 consecutive lines of one plane; lane j solves row A = t - 2 j of line B0 + j at step t.
// Stream (doubles): index = ((step_base(C, s) + t) * NE + e) * 32 + j, NE = 13 (forward) or 14 (backward); pad entries are 0.
#pragma once
#ifndef __CUDACC__
#define SK_HD
#else
#define SK_HD __host__ __device__
#endif

namespace b200 {

struct SkewGeom {
  int NR = 0, NL = 0, NP = 0;     // rows per line, lines per plane, planes
  int BW = 0, S = 0;              // lines per strip (<= 32), strips per plane
  int steps_full = 0;             // steps of a full strip: NR + 2 (BW - 1)
  int steps_plane = 0;            // sum over the strips of a plane
  SK_HD int nb(int s) const { int r = NL - s * BW; return r < BW ? r : BW; }
  SK_HD int nsteps(int s) const { return NR + 2 * (nb(s) - 1); }
  SK_HD long long step_base(int C, int s) const { return (long long)C * steps_plane + (long long)s * steps_full; }
  SK_HD long long ntasks() const { return (long long)NP * S; }
  SK_HD long long total_steps() const { return (long long)NP * steps_plane; }
  SK_HD long long nat(bool upper, int A, int B, int C) const {
    return upper ? (long long)(NR - 1 - A) + (long long)NR * ((NL - 1 - B) + (long long)NL * (NP - 1 - C))
                 : (long long)A + (long long)NR * (B + (long long)NL * C);
  }
  SK_HD bool inside(int A, int B, int C) const { return A >= 0 && A < NR && B >= 0 && B < NL && C >= 0 && C < NP; }
};

// operand offsets (dA, dB, dC) of slot e in sweep coordinates
SK_HD inline void sk_offset(int e, int &dA, int &dB, int &dC) {
  if (e < 9) { dA = e % 3 - 1; dB = e / 3 - 1; dC = -1; }
  else if (e < 12) { dA = e - 10; dB = -1; dC = 0; }
  else { dA = -1; dB = 0; dC = 0; }
}
// slot of an offset, -1 if the offset is not one of the 13 earlier neighbours
SK_HD inline int sk_slot(int dA, int dB, int dC) {
  if (dA < -1 || dA > 1 || dB < -1 || dB > 1) return -1;
  if (dC == -1) return (dA + 1) + 3 * (dB + 1);
  if (dC != 0) return -1;
  if (dB == -1) return 10 + dA;
  if (dB == 0 && dA == -1) return 12;
  return -1;
}

// Scatters row i of the ILU factor (CRS order, inverse diagonal stored on the diagonal, CRSMatrix.F90:3654-3660) into the forward
// stream SL (13 slots per step) and the backward stream SU (14 slots, slot 13 = inverse diagonal).  Used by the device fill
// kernel and by the CPU harness, so both see the same layout.
SK_HD inline void sk_fill_row(const SkewGeom &g, int i, const int *rows, const int *cols, const double *ilu, double *SL, double *SU) {
  const int a = i % g.NR, b = (i / g.NR) % g.NL, c = i / (g.NR * g.NL);
  const int sf = b / g.BW, jf = b - sf * g.BW;                              // forward sweep position of this row
  const long long pf = (g.step_base(c, sf) + a + 2 * jf) * 13;
  const int A = g.NR - 1 - a, B = g.NL - 1 - b, C = g.NP - 1 - c;           // backward sweep position (mirrored coordinates)
  const int sb = B / g.BW, jb = B - sb * g.BW;
  const long long pb = (g.step_base(C, sb) + A + 2 * jb) * 14;
  for (int p = rows[i]; p < rows[i + 1]; ++p) {
    const int j = cols[p];
    const int da = j % g.NR - a, db = (j / g.NR) % g.NL - b, dc = j / (g.NR * g.NL) - c;
    if (j < i) { const int e = sk_slot(da, db, dc); if (e >= 0) SL[(pf + e) * 32 + jf] = ilu[p]; }
    else if (j > i) { const int e = sk_slot(-da, -db, -dc); if (e >= 0) SU[(pb + e) * 32 + jb] = ilu[p]; }
    else SU[(pb + 13) * 32 + jb] = ilu[p];
  }
}

inline void sk_set_strips(SkewGeom &g, int bw_max = 32) {
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

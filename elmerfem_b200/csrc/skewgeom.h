// Grid-stencil detection and operand-slot numbering shared by the tile triangular solves (wave.cu / wavegeom.h, lane.cu / lanegeom.h) and
// their CPU emulations (tests/wave_harness.cpp, tests/lane_harness.cpp).  Plain C++.
//
// The matrix must be the ILU(0) factor of the 27-point stencil on an NR x NL x NP grid in natural numbering, i = a + NR (b + NL c) -- the
// numbering ElmerGrid gives structured meshes; sk_detect finds the three extents from the CRS pattern and verifies EVERY row.
//
// Operand slots e = 0..12 are the 13 earlier neighbours of a row in sweep coordinates,
//   e = (dA+1) + 3 (dB+1) for dC = -1;  e = 10 + dA for dC = 0, dB = -1;  e = 12 for (dA, dB, dC) = (-1, 0, 0):
// ascending natural column order for the forward sweep; for the backward sweep (the same program on the mirrored grid) DESCENDING e is
// ascending column order.  [The file keeps the name of the round-1 "skewed-lane" kernel it was written for; that kernel (8.6 ms per
// application on the 200^3 problem) and its layout helpers were removed in round 2.]
#pragma once
#ifndef __CUDACC__
#define SK_HD
#else
#define SK_HD __host__ __device__
#endif

namespace b200 {

struct SkewGeom { int NR = 0, NL = 0, NP = 0; };   // rows per line, lines per plane, planes

// slot of an offset in sweep coordinates, -1 if the offset is not one of the 13 earlier neighbours
SK_HD inline int sk_slot(int dA, int dB, int dC) {
  if (dA < -1 || dA > 1 || dB < -1 || dB > 1) return -1;
  if (dC == -1) return (dA + 1) + 3 * (dB + 1);
  if (dC != 0) return -1;
  if (dB == -1) return 10 + dA;
  if (dB == 0 && dA == -1) return 12;
  return -1;
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
  return nullptr;
}

}  // namespace b200

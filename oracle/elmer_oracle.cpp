// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU restatement (the "oracle") of Elmer's sparse iterative linear-solve path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; nothing under elmerfem_b200/ links or imports it.
//
// Every routine follows the loop order, guards and constants of the reference file:line
// cited above it (paths relative to the ElmerCSC/elmerfem tree).  The reference is Fortran
// and cannot be compiled in this image (no Fortran compiler), so this restatement is pinned
// against the reference's golden vectors instead (tests/test_oracle_golden.py):
//   fhutiter/examples/ex1/testmat(.out), fem/tests/PoissonThreaded (norm 0.24103925E-01),
//   and method-independence of the answer as in fem/tests/linearsolvers/TempDist.sif.
// Iteration counts are pinned by no reference test: count parity is "unpinned".
//
// Floating point: compiled with -ffp-contract=off (gfortran -O2 on x86-64 without -march
// emits no FMA either), no -ffast-math.  OpenMP only where the reference has !$omp.
//
// All index arrays are the reference's raw 1-based Fortran arrays.

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <climits>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// Types.F90:71  AEPS = 10*EPSILON(1.0_dp)
const double AEPS = 10.0 * DBL_EPSILON;
// huti_fdefs.h:14  HUTI_EPSILON
const double HUTI_EPSILON = 1.17549435E-38;

// huti_fdefs.h:18-50 status codes
enum { HUTI_OK = 0, HUTI_CONVERGENCE = 1, HUTI_MAXITER = 2, HUTI_DIVERGENCE = 3, HUTI_HALTED = 4,
       HUTI_CG_RHO = 20, HUTI_BICGSTAB_RHO = 35, HUTI_BICGSTAB_SNORM = 36, HUTI_BICGSTAB_OMEGA = 37 };
// huti_fdefs.h:85-92 stopping criteria
enum { HUTI_TRUERESIDUAL = 0, HUTI_TRESID_SCALED_BYB = 1, HUTI_PSEUDORESIDUAL = 2,
       HUTI_PRESID_SCALED_BYB = 3, HUTI_PRESID_SCALED_BYPRECB = 4, HUTI_XDIFF_NORM = 5 };

// ipar/dpar slots, huti_fdefs.h:101-155 (1-based Fortran slot k is ipar[k-1])
#define IPAR(k) ipar[(k) - 1]
#define DPAR(k) dpar[(k) - 1]
#define HUTI_NDIM IPAR(3)
#define HUTI_DBUGLVL IPAR(5)
#define HUTI_MAXIT IPAR(10)
#define HUTI_MINIT IPAR(11)
#define HUTI_STOPC IPAR(12)
#define HUTI_GMRES_RESTART IPAR(15)
#define HUTI_BICGSTABL_L IPAR(16)
#define HUTI_GCR_RESTART IPAR(17)
#define HUTI_IDRS_S IPAR(18)
#define HUTI_SMOOTHING IPAR(28)
#define HUTI_INFO IPAR(30)
#define HUTI_ITERS IPAR(31)
#define HUTI_TOLERANCE DPAR(1)
#define HUTI_MAXTOLERANCE DPAR(2)

struct Matrix {
  int n;
  const int *Rows, *Cols, *Diag;      // 1-based contents, 0-based C storage
  const double *Values;
  const double *ILUValues;            // may be null
  int ndeg;
  int precond;                        // 0 none, 1 diagonal, 2 ilu
  long n_matvec, n_pcond, n_dot, n_norm;
  const int *ILURows = nullptr, *ILUCols = nullptr, *ILUDiag = nullptr;   // null: alias Rows/Cols/Diag (ILU0, CRSMatrix.F90:3488-3491)
};

// ---------------------------------------------------------------------------------------
// mathlibs/src/blas/ddot.f (incx=incy=1 branch): single accumulator, mod-5 clean-up loop,
// then unrolled by 5 with the five products added left to right into dtemp.
// g_dot_order != 0 is NOT the reference: it replays the same dot product in the summation orders other BLAS
// builds use (Elmer may be linked against OpenBLAS/MKL instead of mathlibs, CMakeLists.txt:270-...), to
// measure how far iteration counts move with the summation order alone.
//   1: eight interleaved partial sums (SIMD-style kernel), 2: pairwise over blocks of 256, 3: the device's order (below)
static int g_dot_order = 0;
// g_dot_order == 3: the summation order of the DEVICE reductions (elmerfem_b200/csrc/common.cuh grid_reduce, blas1.cu
// k_dot_batch, the SpMV epilogues of spmv.cu and the fused vector kernels of krylov.cu), so that whole solves can be compared
// bit for bit: B = min(ceil(n/256), g_dev_blocks) blocks of 256 threads; thread t adds the products of elements t, t + 256 B, ...
// to one accumulator (separate multiply and add roundings); the 32 lanes of a warp are summed by the xor-shuffle tree
// (offsets 16, 8, 4, 2, 1); the 8 warp sums of a block by the same tree; the block partials are added by 256 threads with stride
// 256, then the same two trees.  The SELL SpMV kernels own rows (slice = warp, warp + 8 B, ...; row = 32 slice + lane), which is the
// same element -> thread map.  Single-rank handles only (an all-reduce adds its own order).
static int g_dev_blocks = 148 * 8;
static inline double tree32(double *v) {            // v[32] is overwritten
  for (int o = 16; o > 0; o >>= 1) for (int i = 0; i < o; ++i) v[i] = v[i] + v[i + o];
  return v[0];
}
static double block256(const double *acc) {          // 256 thread accumulators -> block partial
  double w[32];
  for (int k = 0; k < 32; ++k) w[k] = 0.0;
  for (int warp = 0; warp < 8; ++warp) {
    double v[32];
    for (int l = 0; l < 32; ++l) v[l] = acc[warp * 32 + l];
    w[warp] = tree32(v);
  }
  return tree32(w);
}
static double dot_device(long n, const double *x, const double *y) {
  if (n <= 0) return 0.0;
  long B = (n + 255) / 256;
  if (B > g_dev_blocks) B = g_dev_blocks;
  const long T = B * 256;
  std::vector<double> acc((size_t)T, 0.0);
#pragma omp parallel for schedule(static)
  for (long t = 0; t < T; ++t) {
    double a = 0.0;
    for (long i = t; i < n; i += T) { double p = x[i] * y[i]; a = a + p; }
    acc[t] = a;
  }
  std::vector<double> part((size_t)B);
#pragma omp parallel for schedule(static)
  for (long b = 0; b < B; ++b) part[b] = block256(&acc[(size_t)b * 256]);
  double fin[256];
  for (int t = 0; t < 256; ++t) { double a = 0.0; for (long i = t; i < B; i += 256) a = a + part[i]; fin[t] = a; }
  return block256(fin);
}
static double dot_pairwise(const double *x, const double *y, long lo, long hi) {
  if (hi - lo <= 256) { double s = 0.0; for (long i = lo; i < hi; ++i) s += x[i] * y[i]; return s; }
  long mid = lo + (hi - lo) / 2;
  return dot_pairwise(x, y, lo, mid) + dot_pairwise(x, y, mid, hi);
}
// ddot on the small dense work arrays of BiCGStab(l) (IterativeMethods.F90:960-1037): always the reference's sequential order --
// the device evaluates these in one thread, left to right, so the alternative summation orders do not apply to them.
static double ref_ddot_small(int n, const double *dx, const double *dy) {
  double dtemp = 0.0;
  for (int i = 0; i < n; ++i) dtemp = dtemp + dx[i] * dy[i];       // ddot.f: m = n % 5 first, then groups of five -- all left to right
  return dtemp;
}
double ref_ddot(int n, const double *dx, const double *dy) {
  double dtemp = 0.0;
  if (n <= 0) return 0.0;
  if (g_dot_order == 1) {
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i = 0;
    for (; i + 8 <= n; i += 8) for (int k = 0; k < 8; ++k) a[k] += dx[i + k] * dy[i + k];
    for (; i < n; ++i) a[0] += dx[i] * dy[i];
    return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
  if (g_dot_order == 2) return dot_pairwise(dx, dy, 0, n);
  if (g_dot_order == 3) return dot_device(n, dx, dy);
  int m = n % 5;
  if (m != 0) {
    for (int i = 0; i < m; ++i) dtemp = dtemp + dx[i] * dy[i];
    if (n < 5) return dtemp;
  }
  for (int i = m; i < n; i += 5)
    dtemp = dtemp + dx[i] * dy[i] + dx[i + 1] * dy[i + 1] + dx[i + 2] * dy[i + 2] +
            dx[i + 3] * dy[i + 3] + dx[i + 4] * dy[i + 4];
  return dtemp;
}

// mathlibs/src/blas/dnrm2.f: one-pass scaled sum of squares.
double ref_dnrm2(int n, const double *x) {
  if (n < 1) return 0.0;
  if (n == 1) return std::fabs(x[0]);
  if (g_dot_order != 0) return std::sqrt(ref_ddot(n, x, x));
  double scale = 0.0, ssq = 1.0;
  for (int ix = 0; ix < n; ++ix) {
    if (x[ix] != 0.0) {
      double absxi = std::fabs(x[ix]);
      if (scale < absxi) {
        double q = scale / absxi;
        ssq = 1.0 + ssq * (q * q);
        scale = absxi;
      } else {
        double q = absxi / scale;
        ssq = ssq + q * q;
      }
    }
  }
  return scale * std::sqrt(ssq);
}

// ---------------------------------------------------------------------------------------
// CRSMatrix.F90:4744-4905 CRS_MatrixVectorProd, non-transposed branch, no MKL, no external hook.
// ndeg in {2,3,4,5,6,8,10} selects the variants that read one column index per ndeg entries
// and keep ndeg partial sums (4794-4856); everything else takes the default row loop (4857-4867).
void crs_matvec(const Matrix &A, const double *u, double *v) {
  const int n = A.n;
  const int *Rows = A.Rows, *Cols = A.Cols;
  const double *Values = A.Values;
  const double *U = u - 1;  // 1-based column access
  switch (A.ndeg) {
    case 5: case 10:
#pragma omp parallel for
      for (int i = 0; i < n; ++i) {
        double r1 = 0, r2 = 0, r3 = 0, r4 = 0, r5 = 0;
        for (int j = Rows[i] - 1; j < Rows[i + 1] - 1; j += 5) {
          int l = Cols[j];
          r1 = r1 + U[l] * Values[j];
          r2 = r2 + U[l + 1] * Values[j + 1];
          r3 = r3 + U[l + 2] * Values[j + 2];
          r4 = r4 + U[l + 3] * Values[j + 3];
          r5 = r5 + U[l + 4] * Values[j + 4];
        }
        v[i] = r1 + r2 + r3 + r4 + r5;
      }
      break;
    case 4: case 8:
#pragma omp parallel for
      for (int i = 0; i < n; ++i) {
        double r1 = 0, r2 = 0, r3 = 0, r4 = 0;
        for (int j = Rows[i] - 1; j < Rows[i + 1] - 1; j += 4) {
          int l = Cols[j];
          r1 = r1 + U[l] * Values[j];
          r2 = r2 + U[l + 1] * Values[j + 1];
          r3 = r3 + U[l + 2] * Values[j + 2];
          r4 = r4 + U[l + 3] * Values[j + 3];
        }
        v[i] = r1 + r2 + r3 + r4;
      }
      break;
    case 3: case 6:
#pragma omp parallel for
      for (int i = 0; i < n; ++i) {
        double r1 = 0, r2 = 0, r3 = 0;
        for (int j = Rows[i] - 1; j < Rows[i + 1] - 1; j += 3) {
          int l = Cols[j];
          r1 = r1 + U[l] * Values[j];
          r2 = r2 + U[l + 1] * Values[j + 1];
          r3 = r3 + U[l + 2] * Values[j + 2];
        }
        v[i] = r1 + r2 + r3;
      }
      break;
    case 2:
#pragma omp parallel for
      for (int i = 0; i < n; ++i) {
        double r1 = 0, r2 = 0;
        for (int j = Rows[i] - 1; j < Rows[i + 1] - 1; j += 2) {
          int l = Cols[j];
          r1 = r1 + U[l] * Values[j];
          r2 = r2 + U[l + 1] * Values[j + 1];
        }
        v[i] = r1 + r2;
      }
      break;
    default:
#pragma omp parallel for
      for (int i = 0; i < n; ++i) {
        double r1 = 0.0;
        for (int j = Rows[i] - 1; j < Rows[i + 1] - 1; ++j) r1 = r1 + U[Cols[j]] * Values[j];
        v[i] = r1;
      }
  }
}

// CRSMatrix.F90:2279-2326 CRS_DiagPrecondition (matrix already Ordered, Diag filled).
void crs_diag_precond(const Matrix &A, double *u, const double *v) {
  const int n = A.n;
#pragma omp parallel for
  for (int i = 0; i < n; ++i) {
    double d = A.Values[A.Diag[i] - 1];
    if (std::fabs(d) > AEPS) u[i] = v[i] / d;
    else u[i] = v[i];
  }
}

// CRSMatrix.F90:3445-3531 + 3604-3661 CRS_IncompleteLU, ILUn == 0, non-Cholesky branch.
// ILURows/ILUCols/ILUDiag alias Rows/Cols/Diag (3488-3491).  Serial, as in the reference.
int crs_ilu0(int N, const int *Rows, const int *Cols, const int *Diag, const double *Values,
             double *ILUValues) {
  if (N == 0) return 1;
  std::vector<char> C(N + 1, 0);
  std::vector<double> S(N + 1, 0.0);
  for (int i = 1; i <= N; ++i) {
    // 3614-3620: scatter the row to full form, flag the pattern
    for (int k = Rows[i - 1]; k <= Rows[i] - 1; ++k) S[Cols[k - 1]] = Values[k - 1];
    for (int k = Rows[i - 1]; k <= Rows[i] - 1; ++k) C[Cols[k - 1]] = 1;
    // 3624-3637: eliminate with the finished rows k < i of the pattern, in column order
    for (int m = Rows[i - 1]; m <= Diag[i - 1] - 1; ++m) {
      int k = Cols[m - 1];
      if (S[k] == 0.0) continue;
      double ukk = ILUValues[Diag[k - 1] - 1];
      if (std::fabs(ukk) > AEPS) S[k] = S[k] / ukk;
      for (int l = Diag[k - 1] + 1; l <= Rows[k] - 1; ++l) {
        int j = Cols[l - 1];
        if (C[j]) S[j] = S[j] - S[k] * ILUValues[l - 1];
      }
    }
    // 3643-3649: gather back
    for (int k = Rows[i - 1]; k <= Rows[i] - 1; ++k) {
      int c = Cols[k - 1];
      if (C[c]) {
        ILUValues[k - 1] = S[c];
        S[c] = 0.0;
        C[c] = 0;
      }
    }
  }
  // 3654-3660: prescale (invert) the diagonal for the LU solve
  for (int i = 1; i <= N; ++i) {
    double &d = ILUValues[Diag[i - 1] - 1];
    if (std::fabs(d) < AEPS) d = 1.0;
    else d = 1.0 / d;
  }
  return 1;
}

// CRSMatrix.F90:3664-3795 InitializeILU1: the pattern of one more level of fill.  Row i keeps its entries and gains
// the columns j of the upper parts of the rows k < i it already holds (only the rows flagged 1, i.e. present BEFORE
// this round: fills created in the round do not cascade).  Columns come out ascending.  Two calls: ILUCols == null
// counts (3694-3717), otherwise fills (3729-3762).  1-based contents.
long crs_ilu1_pattern(int N, const int *Rows, const int *Cols, const int *Diag, int *ILURows, int *ILUCols, int *ILUDiag) {
  std::vector<int> C(N + 2, 0);
  long nz = 0;
  if (ILURows) ILURows[0] = 1;
  for (int i = 1; i <= N; ++i) {
    for (int k = Rows[i - 1]; k <= Rows[i] - 1; ++k) C[Cols[k - 1]] = 1;
    int RowMin = Cols[Rows[i - 1] - 1], RowMax = Cols[Rows[i] - 2];
    for (int k = RowMin; k <= i - 1; ++k) {
      if (C[k] == 1) {
        for (int l = Diag[k - 1] + 1; l <= Rows[k] - 1; ++l) {
          int j = Cols[l - 1];
          if (C[j] == 0) { C[j] = 2; RowMax = std::max(RowMax, j); }
        }
      }
    }
    long j = ILURows ? ILURows[i - 1] - 1 : 0;
    for (int k = RowMin; k <= RowMax; ++k) {
      if (C[k] > 0) {
        ++j; ++nz;
        C[k] = 0;
        if (ILUCols) { ILUCols[j - 1] = k; if (k == i) ILUDiag[i - 1] = (int)j; }
      }
    }
    if (ILURows) ILURows[i] = (int)(j + 1);
  }
  return nz;
}

// CRSMatrix.F90:3604-3661 with ILUn > 0: the same row-by-row elimination on the ILU pattern; the row is scattered
// from A's own entries (fill positions start from 0) and only pattern positions are updated.
int crs_ilun_factor(int N, const int *Rows, const int *Cols, const double *Values, const int *ILURows, const int *ILUCols,
                    const int *ILUDiag, double *ILUValues) {
  if (N == 0) return 1;
  std::vector<char> C(N + 1, 0);
  std::vector<double> S(N + 1, 0.0);
  for (int i = 1; i <= N; ++i) {
    for (int k = Rows[i - 1]; k <= Rows[i] - 1; ++k) S[Cols[k - 1]] = Values[k - 1];
    for (int k = ILURows[i - 1]; k <= ILURows[i] - 1; ++k) C[ILUCols[k - 1]] = 1;
    for (int m = ILURows[i - 1]; m <= ILUDiag[i - 1] - 1; ++m) {
      int k = ILUCols[m - 1];
      if (S[k] == 0.0) continue;
      double ukk = ILUValues[ILUDiag[k - 1] - 1];
      if (std::fabs(ukk) > AEPS) S[k] = S[k] / ukk;
      for (int l = ILUDiag[k - 1] + 1; l <= ILURows[k] - 1; ++l) {
        int j = ILUCols[l - 1];
        if (C[j]) S[j] = S[j] - S[k] * ILUValues[l - 1];
      }
    }
    for (int k = ILURows[i - 1]; k <= ILURows[i] - 1; ++k) {
      int c = ILUCols[k - 1];
      if (C[c]) { ILUValues[k - 1] = S[c]; S[c] = 0.0; C[c] = 0; }
    }
  }
  for (int i = 1; i <= N; ++i) {
    double &d = ILUValues[ILUDiag[i - 1] - 1];
    if (std::fabs(d) < AEPS) d = 1.0;
    else d = 1.0 / d;
  }
  return 1;
}

// CRSMatrix.F90:3539-3602: the Cholesky branch of CRS_IncompleteLU (A % Cholesky, set from 'Linear System Symmetric ILU',
// IterSolve.F90:526).  Row by row: T = row i of the matrix (lower part and diagonal) in full form; for every lower entry j of the ILU
// pattern, in column order, S(j) = (T(j) - sum_l S(k_l) L(j, k_l)) * L(j, j)^-1 over the WHOLE lower part of row j (S is zero outside
// row i's pattern), and S(i) = T(i) - sum_j S(j)^2; the diagonal is stored as 1 / sqrt(S(i)) (1 when S(i) <= AEPS).  Only the lower
// part and the diagonal of ILUValues are written (the reference leaves the rest of the freshly allocated array untouched; here 0).
int crs_ichol_factor(int N, const int *Rows, const int *Cols, const int *Diag, const double *Values, const int *ILURows, const int *ILUCols,
                     const int *ILUDiag, double *ILUValues) {
  if (N == 0) return 1;
  std::vector<double> S(N + 1, 0.0), T(N + 1, 0.0);
  for (long q = 0; q < (long)ILURows[N] - 1; ++q) ILUValues[q] = 0.0;
  for (int i = 1; i <= N; ++i) {
    for (int k = Rows[i - 1]; k <= Diag[i - 1]; ++k) T[Cols[k - 1]] = Values[k - 1];                     // 3553-3556
    for (int k = ILURows[i - 1]; k <= ILUDiag[i - 1]; ++k) S[ILUCols[k - 1]] = ILUValues[k - 1];          // 3558-3562
    S[i] = T[i];                                                                                          // 3566
    for (int m = ILURows[i - 1]; m <= ILUDiag[i - 1] - 1; ++m) {                                          // 3567-3576
      const int j = ILUCols[m - 1];
      S[j] = T[j];
      for (int l = ILURows[j - 1]; l <= ILUDiag[j - 1] - 1; ++l) {
        const int k = ILUCols[l - 1];
        S[j] = S[j] - S[k] * ILUValues[l - 1];
      }
      S[j] = S[j] * ILUValues[ILUDiag[j - 1] - 1];
      S[i] = S[i] - S[j] * S[j];
    }
    if (S[i] <= AEPS) S[i] = 1.0;                                                                         // 3578-3587
    else S[i] = 1.0 / std::sqrt(S[i]);
    for (int k = Rows[i - 1]; k <= Diag[i - 1]; ++k) T[Cols[k - 1]] = 0.0;                                // 3591-3594
    for (int k = ILURows[i - 1]; k <= ILUDiag[i - 1]; ++k) {                                              // 3596-3601
      const int j = ILUCols[k - 1];
      ILUValues[k - 1] = S[j];
      S[j] = 0.0;
    }
  }
  return 1;
}

// CRSMatrix.F90:4144-4340 CRS_ILUT / ComputeILUT: incomplete LU with threshold dropping.  Row by row: the row is scattered into full form,
// eliminated against every flagged column k < i in ascending order (fill-ins of the upper parts of the pivot rows join the flagged set as
// they appear), then every flagged entry with |S(k)| >= TOL * ||A(i,:)||_2 -- and always the diagonal -- is stored in column order; the
// diagonal is inverted at the end (4323-4329).  The pattern is an OUTPUT.  Returns the number of stored entries, or -(needed so far) when
// `cap` entries do not suffice.  1-based contents.
long crs_ilut(int N, const int *Rows, const int *Cols, const double *Values, double TOL, long cap, int *ILURows, int *ILUCols, int *ILUDiag,
              double *ILUValues) {
  std::vector<char> C(N + 2, 0);
  std::vector<double> S(N + 2, 0.0);
  ILURows[0] = 1;
  for (int i = 1; i <= N; ++i) {
    for (int k = Rows[i - 1]; k <= Rows[i] - 1; ++k) { C[Cols[k - 1]] = 1; S[Cols[k - 1]] = Values[k - 1]; }
    int RowMin = Cols[Rows[i - 1] - 1], RowMax = Cols[Rows[i] - 2];
    for (int k = RowMin; k <= i - 1; ++k) {
      if (C[k]) {
        const double piv = ILUValues[ILUDiag[k - 1] - 1];
        if (std::fabs(piv) > AEPS) S[k] = S[k] / piv;
        for (int l = ILUDiag[k - 1] + 1; l <= ILURows[k] - 1; ++l) {
          const int j = ILUCols[l - 1];
          if (!C[j]) { C[j] = 1; RowMax = std::max(RowMax, j); }
          S[j] = S[j] - S[k] * ILUValues[l - 1];
        }
      }
    }
    double s2 = 0.0;                                                    // 4262: SQRT( SUM( ABS(Values(row))**2 ) ), summed in order
    for (int k = Rows[i - 1]; k <= Rows[i] - 1; ++k) { const double a = std::fabs(Values[k - 1]); s2 = s2 + a * a; }
    const double NORMA = std::sqrt(s2);
    long j = (long)ILURows[i - 1] - 1;
    for (int k = RowMin; k <= RowMax; ++k) {
      if (C[k]) {
        if (std::fabs(S[k]) >= TOL * NORMA || k == i) {
          ++j;
          if (j > cap) return -j;
          ILUCols[j - 1] = k; ILUValues[j - 1] = S[k];
          if (k == i) ILUDiag[i - 1] = (int)j;
        }
        S[k] = 0.0; C[k] = 0;
      }
    }
    ILURows[i] = (int)(j + 1);
  }
  for (int i = 1; i <= N; ++i) {
    double &d = ILUValues[ILUDiag[i - 1] - 1];
    if (std::fabs(d) < AEPS) d = 1.0;
    else d = 1.0 / d;
  }
  return (long)ILURows[N] - 1;
}

// A % Cholesky of the matrix the callbacks below work on ('Linear System Symmetric ILU')
static int g_cholesky = 0;

// CRSMatrix.F90:4590-4663 CRS_LUSolve: Cholesky branch 4618-4638, LU branch 4642-4660; diagonal fallback 4610-4616.
void crs_lusolve(const Matrix &A, double *b) {
  const int n = A.n;
  const int *Rows = A.ILURows ? A.ILURows : A.Rows, *Cols = A.ILUCols ? A.ILUCols : A.Cols, *Diag = A.ILUDiag ? A.ILUDiag : A.Diag;
  const double *Values = A.ILUValues;
  double *B = b - 1;
  if (!Values) {
    for (int i = 1; i <= n; ++i) {
      double s = A.Values[A.Diag[i - 1] - 1];
      if (s != 0) B[i] = B[i] / s;
    }
    return;
  }
  if (g_cholesky) {
    for (int i = 1; i <= n; ++i) {                                   // 4621-4628: L z = b
      double s = B[i];
      for (int j = Rows[i - 1]; j <= Diag[i - 1] - 1; ++j) s = s - Values[j - 1] * B[Cols[j - 1]];
      B[i] = s * Values[Diag[i - 1] - 1];
    }
    for (int i = n; i >= 1; --i) {                                   // 4632-4638: L^T x = z, column-oriented
      B[i] = B[i] * Values[Diag[i - 1] - 1];
      for (int j = Rows[i - 1]; j <= Diag[i - 1] - 1; ++j) B[Cols[j - 1]] = B[Cols[j - 1]] - Values[j - 1] * B[i];
    }
    return;
  }
  for (int i = 1; i <= n; ++i) {
    double s = B[i];
    for (int j = Rows[i - 1]; j <= Diag[i - 1] - 1; ++j) s = s - Values[j - 1] * B[Cols[j - 1]];
    B[i] = s;
  }
  for (int i = n; i >= 1; --i) {
    double s = B[i];
    for (int j = Diag[i - 1] + 1; j <= Rows[i] - 1; ++j) s = s - Values[j - 1] * B[Cols[j - 1]];
    B[i] = Values[Diag[i - 1] - 1] * s;
  }
}

// CRSMatrix.F90:4550-4564 CRS_LUPrecondition: copy (OMP) then in-place solve.
void crs_lu_precond(const Matrix &A, double *u, const double *v) {
  const int n = A.n;
#pragma omp parallel for
  for (int i = 0; i < n; ++i) u[i] = v[i];
  crs_lusolve(A, u);
}

// IterSolve.F90:121-133 pcond_dummy / huti_aux.F90:167-185 huti_ddummy_pcondfun: u = v.
void pcond_dummy(const Matrix &A, double *u, const double *v) {
  for (int i = 0; i < A.n; ++i) u[i] = v[i];
}

// --- the five HUTI callbacks bound to the matrix (IterSolve.F90:790-912) ---
struct Ops {
  Matrix *A;
  bool left = false;   // IterSolve.F90:509-525: GMRES (BiCGStab2, TFQMR) get the preconditioner in the LEFT slot, the dummy in the right one
  void matvec(const double *u, double *v) { crs_matvec(*A, u, v); A->n_matvec++; }
  void apply(double *u, const double *v) {
    A->n_pcond++;
    if (A->precond == 2) crs_lu_precond(*A, u, v);
    else if (A->precond == 1) crs_diag_precond(*A, u, v);
    else pcond_dummy(*A, u, v);
  }
  void pcondr(double *u, const double *v) { if (left) pcond_dummy(*A, u, v); else apply(u, v); }   // right preconditioner slot
  void pcondl(double *u, const double *v) { if (left) apply(u, v); else pcond_dummy(*A, u, v); }   // pconddProc = 0 -> dummy
  double dot(int n, const double *x, const double *y) { A->n_dot++; return ref_ddot(n, x, y); }
  double norm(int n, const double *x) { A->n_norm++; return ref_dnrm2(n, x); }
};

// ---------------------------------------------------------------------------------------
// fhutiter/src/huti_cg.F90:267-518 huti_dcgsolv.  work(n,4) = Z,P,Q,R (huti_cg.F90 macros).
void huti_dcgsolv(Ops &op, int ndim, double *X, const double *B, int *ipar, double *dpar,
                  double *work) {
  double *Z = work, *P = work + (size_t)ndim, *Q = work + 2 * (size_t)ndim, *R = work + 3 * (size_t)ndim;
  double alpha = 0, beta = 0, rho = 0, oldrho = 0, residual = 0, rhsnorm = 1.0;
  int iter_count = 1;
  if (HUTI_STOPC == HUTI_TRESID_SCALED_BYB || HUTI_STOPC == HUTI_PRESID_SCALED_BYB)
    rhsnorm = op.norm(ndim, B);
  op.matvec(X, R);
#pragma omp parallel for
  for (int i = 0; i < ndim; ++i) R[i] = B[i] - R[i];
  for (;;) {
    op.pcondl(Q, R);
    op.pcondr(Z, Q);
    rho = op.dot(ndim, R, Z);
    if (rho == 0) { HUTI_INFO = HUTI_CG_RHO; break; }
    if (iter_count == 1) {
#pragma omp parallel for
      for (int i = 0; i < ndim; ++i) P[i] = Z[i];
    } else {
      beta = rho / oldrho;
#pragma omp parallel for
      for (int i = 0; i < ndim; ++i) P[i] = Z[i] + beta * P[i];
    }
    op.matvec(P, Q);
    alpha = rho / op.dot(ndim, P, Q);
#pragma omp parallel for
    for (int i = 0; i < ndim; ++i) X[i] = X[i] + alpha * P[i];
#pragma omp parallel for
    for (int i = 0; i < ndim; ++i) R[i] = R[i] - alpha * Q[i];
    switch (HUTI_STOPC) {
      case HUTI_TRESID_SCALED_BYB:
        op.matvec(X, Z);
#pragma omp parallel for
        for (int i = 0; i < ndim; ++i) Z[i] = Z[i] - B[i];
        residual = op.norm(ndim, Z) / rhsnorm;
        break;
      case HUTI_PSEUDORESIDUAL: residual = op.norm(ndim, R); break;
      case HUTI_PRESID_SCALED_BYB: residual = op.norm(ndim, R) / rhsnorm; break;
      case HUTI_XDIFF_NORM:
        for (int i = 0; i < ndim; ++i) Z[i] = alpha * P[i];
        residual = op.norm(ndim, Z);
        break;
      default:  // HUTI_TRUERESIDUAL and "case default"
        op.matvec(X, Z);
#pragma omp parallel for
        for (int i = 0; i < ndim; ++i) Z[i] = Z[i] - B[i];
        residual = op.norm(ndim, Z);
    }
    if (HUTI_DBUGLVL != 0 && HUTI_DBUGLVL != INT_MAX && iter_count % HUTI_DBUGLVL == 0)
      printf("%8d%11.4E\n", iter_count, residual);
    if (residual < HUTI_TOLERANCE) { HUTI_INFO = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > HUTI_MAXTOLERANCE) { HUTI_INFO = HUTI_DIVERGENCE; break; }
    oldrho = rho;
    iter_count = iter_count + 1;
    if (iter_count > HUTI_MAXIT) { HUTI_INFO = HUTI_MAXITER; break; }
  }
  HUTI_ITERS = iter_count;
  dpar[9] = residual;  // test aid only: last residual in dpar(10) (slot unused by HUTI)
}

// fhutiter/src/huti_aux.F90:221-289 huti_dlusolve: in-place LU without pivoting (Saad, Alg. 10.4) of the
// column-major n x n matrix, then L u = v, U u = u.
static void huti_dlusolve(int n, double *lumat, double *u, const double *v) {
#define LU_(i, j) lumat[((i) - 1) + (size_t)((j) - 1) * n]
  for (int i = 2; i <= n; ++i)
    for (int k = 1; k <= i - 1; ++k) {
      LU_(i, k) = LU_(i, k) / LU_(k, k);
      for (int j = k + 1; j <= n; ++j) LU_(i, j) = LU_(i, j) - LU_(i, k) * LU_(k, j);
    }
  for (int i = 1; i <= n; ++i) {
    u[i - 1] = v[i - 1];
    for (int k = 1; k <= i - 1; ++k) u[i - 1] = u[i - 1] - LU_(i, k) * u[k - 1];
  }
  for (int i = n; i >= 1; --i) {
    for (int k = i + 1; k <= n; ++k) u[i - 1] = u[i - 1] - LU_(i, k) * u[k - 1];
    u[i - 1] = u[i - 1] / LU_(i, i);
  }
#undef LU_
}

// fhutiter/src/huti_gmres.F90:390-822 huti_dgmressolv: restarted GMRES(m) with Givens rotations.  work(n, 7+m) =
// W, R, S, VTMP, T1V, V(1..m+1) (macros 59-69).  S keeps the rotated right-hand side in its first m+1 entries.
// In Elmer the preconditioner sits in the left slot (IterSolve.F90:509-525): every residual below is M^-1 (b - A x).
void huti_dgmressolv(Ops &op, int ndim, double *X, const double *B, int *ipar, double *dpar, double *work) {
  const size_t N = (size_t)ndim;
  const int m = HUTI_GMRES_RESTART;
  double *W = work, *R = work + N, *S = work + 2 * N, *VTMP = work + 3 * N, *T1V = work + 4 * N;
  auto V = [&](int i) { return work + (size_t)(5 + i - 1) * N; };      // V(i), i = 1..m+1
  std::vector<double> H((size_t)(m + 1) * (m + 1), 0.0), HLU((size_t)(m + 1) * (m + 1), 0.0), CS(m + 1, 0.0), SN(m + 1, 0.0), Y(m + 1, 0.0);
#define H_(i, j) H[((i) - 1) + (size_t)((j) - 1) * (m + 1)]
  int iter_count = 1;
  double residual = 0, rhsnorm = 1.0, precrhsnorm = 1.0;
  const double bnrm = op.norm(ndim, B);
  if (HUTI_STOPC == HUTI_TRESID_SCALED_BYB || HUTI_STOPC == HUTI_PRESID_SCALED_BYB) rhsnorm = bnrm;
  if (HUTI_STOPC == HUTI_PRESID_SCALED_BYPRECB) { op.pcondl(T1V, B); precrhsnorm = op.norm(ndim, T1V); }
  op.pcondr(T1V, X);                                                 // 475-486 (result overwritten at 300)
  op.matvec(T1V, R);
#pragma omp parallel for
  for (int i = 0; i < ndim; ++i) T1V[i] = B[i] - R[i];
  op.pcondl(R, T1V);
  for (int j = 1; j <= m + 1; ++j) std::fill(V(j), V(j) + N, 0.0);
  std::fill(VTMP, VTMP + N, 0.0);
  VTMP[0] = 1.0;
  for (;;) {                                                          // label 300
    op.pcondr(T1V, X);
    op.matvec(T1V, R);
#pragma omp parallel for
    for (int i = 0; i < ndim; ++i) T1V[i] = B[i] - R[i];
    op.pcondl(R, T1V);
    const double alpha = op.norm(ndim, R);
    if (alpha == 0) { HUTI_INFO = 40; break; }                        // HUTI_GMRES_ALPHA
#pragma omp parallel for
    for (int i = 0; i < ndim; ++i) V(1)[i] = R[i] / alpha;
#pragma omp parallel for
    for (int i = 0; i < ndim; ++i) S[i] = alpha * VTMP[i];
    bool early = false, broke = false;
    for (int i = 1; i <= m; ++i) {
      op.pcondr(W, V(i));
      op.matvec(W, T1V);
      op.pcondl(W, T1V);
      for (int k = 1; k <= i; ++k) {
        H_(k, i) = op.dot(ndim, W, V(k));
        const double hki = H_(k, i); const double *vk = V(k);
#pragma omp parallel for
        for (int ii = 0; ii < ndim; ++ii) W[ii] = W[ii] - hki * vk[ii];
      }
      const double beta = op.norm(ndim, W);
      if (beta == 0) { HUTI_INFO = 41; broke = true; break; }        // HUTI_GMRES_BETA
      H_(i + 1, i) = beta;
      { double *vn = V(i + 1);
#pragma omp parallel for
        for (int ii = 0; ii < ndim; ++ii) vn[ii] = W[ii] / beta; }
      for (int k = 1; k <= i - 1; ++k) {                              // 600-604
        const double temp = CS[k] * H_(k, i) + SN[k] * H_(k + 1, i);
        H_(k + 1, i) = -1 * SN[k] * H_(k, i) + CS[k] * H_(k + 1, i);
        H_(k, i) = temp;
      }
      if (H_(i + 1, i) == 0) { CS[i] = 1; SN[i] = 0; }
      else if (std::fabs(H_(i + 1, i)) > std::fabs(H_(i, i))) {
        const double t2 = H_(i, i) / H_(i + 1, i);
        SN[i] = 1 / std::sqrt(1 + (t2 * t2)); CS[i] = t2 * SN[i];
      } else {
        const double t2 = H_(i + 1, i) / H_(i, i);
        CS[i] = 1 / std::sqrt(1 + (t2 * t2)); SN[i] = t2 * CS[i];
      }
      const double temp = CS[i] * S[i - 1];
      S[i] = -1 * SN[i] * S[i - 1];
      S[i - 1] = temp;
      H_(i, i) = (CS[i] * H_(i, i)) + (SN[i] * H_(i + 1, i));
      H_(i + 1, i) = 0;
      const double error = std::fabs(S[i]) / bnrm;
      if ((float)error < HUTI_TOLERANCE) {                            // 628: REAL(error)
        std::fill(HLU.begin(), HLU.end(), 0.0);
        { size_t j = 0; for (int k = 1; k <= i; ++k) for (int l = 1; l <= i; ++l) HLU[j++] = H_(l, k); }
        huti_dlusolve(i, HLU.data(), Y.data(), S);
        for (int c = 1; c <= i; ++c) { const double y = Y[c - 1]; const double *vc = V(c);
#pragma omp parallel for
          for (int ii = 0; ii < ndim; ++ii) X[ii] = X[ii] + vc[ii] * y; }
        early = true;
        break;
      }
    }
    if (broke) break;
    if (!early) {                                                     // 650-664 (the test at 650 repeats the last `error`)
      std::fill(HLU.begin(), HLU.end(), 0.0);
      { size_t j = 0; for (int k = 1; k <= m; ++k) for (int l = 1; l <= m; ++l) HLU[j++] = H_(l, k); }
      huti_dlusolve(m, HLU.data(), Y.data(), S);
      for (int c = 1; c <= m; ++c) { const double y = Y[c - 1]; const double *vc = V(c);
#pragma omp parallel for
        for (int ii = 0; ii < ndim; ++ii) X[ii] = X[ii] + vc[ii] * y; }
    }
    // 500: convergence check (HUTI_TRESID_SCALED_BYB is what IterSolver sets; the other criteria are restated for completeness)
    if (HUTI_STOPC == HUTI_PSEUDORESIDUAL || HUTI_STOPC == HUTI_PRESID_SCALED_BYB || HUTI_STOPC == HUTI_PRESID_SCALED_BYPRECB) {
      op.matvec(X, R);
#pragma omp parallel for
      for (int i = 0; i < ndim; ++i) R[i] = R[i] - B[i];
      op.pcondl(T1V, R);
      residual = op.norm(ndim, T1V);
      if (HUTI_STOPC == HUTI_PRESID_SCALED_BYB) residual /= rhsnorm;
      if (HUTI_STOPC == HUTI_PRESID_SCALED_BYPRECB) residual /= precrhsnorm;
    } else {
      op.pcondr(T1V, X);
      op.matvec(T1V, R);
#pragma omp parallel for
      for (int i = 0; i < ndim; ++i) T1V[i] = B[i] - R[i];
      op.pcondl(R, T1V);
      residual = op.norm(ndim, R);
      if (HUTI_STOPC == HUTI_TRESID_SCALED_BYB) residual /= rhsnorm;
    }
    S[m] = op.norm(ndim, R);                                          // 781
    if (HUTI_DBUGLVL != 0 && HUTI_DBUGLVL != INT_MAX && iter_count % HUTI_DBUGLVL == 0)
      printf("   gmres:%8d%11.4E\n", iter_count, residual);
    if (residual < HUTI_TOLERANCE) { HUTI_INFO = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > HUTI_MAXTOLERANCE) { HUTI_INFO = HUTI_DIVERGENCE; break; }
    iter_count = iter_count + 1;
    if (iter_count > HUTI_MAXIT) { HUTI_INFO = HUTI_MAXITER; break; }
  }
  HUTI_ITERS = iter_count;
  op.pcondr(T1V, X);                                                  // 815-816
  for (int i = 0; i < ndim; ++i) X[i] = T1V[i];
  dpar[9] = residual;
#undef H_
}

// IterativeMethods.F90:336-391 Jacobi and 444-519 Richardson (preconditioned with the lumped matrix): stationary iterations, no
// preconditioner callback.  `iters` = rounds performed.
static void StationaryIteration(Ops &op, int n, double *x, const double *b, int Rounds, double MinTol, double MaxTol, double &Residual,
                                bool &Converged, bool &Diverged, bool richardson, int *iters) {
  const Matrix &A = *op.A;
  std::vector<double> r(n), M(n);
  Converged = Diverged = false; *iters = 0;
  op.matvec(x, r.data());
  for (int i = 0; i < n; ++i) r[i] = b[i] - r[i];
  const double bnorm = op.norm(n, b);
  double rnorm = op.norm(n, r.data());
  Residual = rnorm / bnorm;
  Converged = Residual < MinTol;
  Diverged = (Residual > MaxTol) || (Residual != Residual);
  if (Converged || Diverged) return;
  if (richardson)
    for (int i = 0; i < n; ++i) { double s = 0.0; for (int j = A.Rows[i] - 1; j < A.Rows[i + 1] - 1; ++j) s = s + A.Values[j]; M[i] = s; }
  for (int k = 1; k <= Rounds; ++k) {
    *iters = k;
    if (richardson) { for (int i = 0; i < n; ++i) x[i] = (k == 1) ? b[i] / M[i] : x[i] + r[i] / M[i]; }
    else for (int j = 0; j < n; ++j) x[j] = x[j] + r[j] / A.Values[A.Diag[j] - 1];
    op.matvec(x, r.data());
    for (int i = 0; i < n; ++i) r[i] = b[i] - r[i];
    rnorm = op.norm(n, r.data());
    Residual = rnorm / bnorm;
    Converged = Residual < MinTol;
    Diverged = (Residual > MaxTol) || (Residual != Residual);
    if (Converged || Diverged) break;
  }
}

// IterativeMethods.F90:219-283 SGS.
static void SGS(Ops &op, int n, double *x, const double *b, int Rounds, double MinTol, double MaxTol, double &Residual, bool &Converged,
                bool &Diverged, double Omega, int *iters) {
  const Matrix &A = *op.A;
  const int *Rows = A.Rows, *Cols = A.Cols; const double *Values = A.Values;
  std::vector<double> r(n);
  Converged = Diverged = false; *iters = 0;
  op.matvec(x, r.data());
  for (int i = 0; i < n; ++i) r[i] = b[i] - r[i];
  const double bnorm = op.norm(n, b);
  double rnorm = op.norm(n, r.data());
  Residual = rnorm / bnorm;
  Converged = Residual < MinTol;
  Diverged = (Residual > MaxTol) || (Residual != Residual);
  if (Converged || Diverged) return;
  for (int k = 1; k <= Rounds; ++k) {
    *iters = k;
    for (int i = 1; i <= n; ++i) {
      double s = 0.0;
      for (int j = Rows[i - 1]; j <= Rows[i] - 1; ++j) s = s + x[Cols[j - 1] - 1] * Values[j - 1];
      x[i - 1] = x[i - 1] + Omega * (b[i - 1] - s) / Values[A.Diag[i - 1] - 1];
    }
    for (int i = n; i >= 1; --i) {
      double s = 0.0;
      for (int j = Rows[i - 1]; j <= Rows[i] - 1; ++j) s = s + x[Cols[j - 1] - 1] * Values[j - 1];
      x[i - 1] = x[i - 1] + Omega * (b[i - 1] - s) / Values[A.Diag[i - 1] - 1];
    }
    op.matvec(x, r.data());
    for (int i = 0; i < n; ++i) r[i] = b[i] - r[i];
    rnorm = op.norm(n, r.data());
    Residual = rnorm / bnorm;
    Converged = Residual < MinTol;
    Diverged = (Residual > MaxTol) || (Residual != Residual);
    if (Converged || Diverged) return;
  }
}

// fhutiter/src/huti_cgs.F90:283-470 huti_dcgssolv.  work(n,7) = RTLD,P,Q,U,T1V,T2V,R.  Right-oriented preconditioning.
void huti_dcgssolv(Ops &op, int ndim, double *X, const double *B, int *ipar, double *dpar, double *work) {
  const size_t N = (size_t)ndim;
  double *RTLD = work, *P = work + N, *Q = work + 2 * N, *U = work + 3 * N, *T1V = work + 4 * N, *T2V = work + 5 * N, *R = work + 6 * N;
  double rho = 0, oldrho = 0, alpha = 0, beta = 0, residual = 0, rhsnorm = 1.0;
  int iter_count = 1;
  if (HUTI_STOPC == HUTI_TRESID_SCALED_BYB || HUTI_STOPC == HUTI_PRESID_SCALED_BYB) rhsnorm = op.norm(ndim, B);
  op.matvec(X, R);
  for (int i = 0; i < ndim; ++i) { R[i] = B[i] - R[i]; RTLD[i] = R[i]; }
  for (;;) {
    rho = op.dot(ndim, RTLD, R);
    if (rho == 0) { HUTI_INFO = 25; break; }                        // HUTI_CGS_RHO
    if (iter_count == 1) {
      for (int i = 0; i < ndim; ++i) { U[i] = R[i]; P[i] = U[i]; }
    } else {
      beta = rho / oldrho;
      for (int i = 0; i < ndim; ++i) U[i] = R[i] + beta * Q[i];
      for (int i = 0; i < ndim; ++i) P[i] = U[i] + beta * Q[i] + beta * beta * P[i];
    }
    op.pcondl(T2V, P); op.pcondr(T1V, T2V);
    op.matvec(T1V, T2V);
    alpha = rho / op.dot(ndim, RTLD, T2V);
    for (int i = 0; i < ndim; ++i) Q[i] = U[i] - alpha * T2V[i];
    for (int i = 0; i < ndim; ++i) T2V[i] = U[i] + Q[i];
    op.pcondl(U, T2V); op.pcondr(T1V, U);
    for (int i = 0; i < ndim; ++i) X[i] = X[i] + alpha * T1V[i];
    op.matvec(T1V, T2V);
    for (int i = 0; i < ndim; ++i) R[i] = R[i] - alpha * T2V[i];
    if (HUTI_STOPC == HUTI_PSEUDORESIDUAL) residual = op.norm(ndim, R);
    else if (HUTI_STOPC == HUTI_PRESID_SCALED_BYB) residual = op.norm(ndim, R) / rhsnorm;
    else {
      op.matvec(X, T1V);
      for (int i = 0; i < ndim; ++i) T1V[i] = T1V[i] - B[i];
      residual = op.norm(ndim, T1V);
      if (HUTI_STOPC == HUTI_TRESID_SCALED_BYB) residual /= rhsnorm;
    }
    if (residual < HUTI_TOLERANCE) { HUTI_INFO = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > HUTI_MAXTOLERANCE) { HUTI_INFO = HUTI_DIVERGENCE; break; }
    oldrho = rho;
    iter_count = iter_count + 1;
    if (iter_count > HUTI_MAXIT) { HUTI_INFO = HUTI_MAXITER; break; }
  }
  HUTI_ITERS = iter_count;
  dpar[9] = residual;
}

// fhutiter/src/huti_tfqmr.F90:455-803 huti_dtfqmrsolv.  work(n,10) = V,Y,YNEW,RTLD,T1V,T2V,W,D,R,TRV.  Elmer puts the
// preconditioner in the left slot (IterSolve.F90:509-525).  Two half steps per iteration, each with its own stopping test.
void huti_dtfqmrsolv(Ops &op, int ndim, double *X, const double *B, int *ipar, double *dpar, double *work) {
  const size_t N = (size_t)ndim;
  double *V = work, *Y = work + N, *YNEW = work + 2 * N, *RTLD = work + 3 * N, *T1V = work + 4 * N, *T2V = work + 5 * N, *W = work + 6 * N,
         *D = work + 7 * N, *R = work + 8 * N, *TRV = work + 9 * N;
  double rho = 0, oldrho = 0, eta = 0, tau = 0, gamma = 0, oldgamma = 0, alpha = 0, beta = 0, c = 0, residual = 0, rhsnorm = 1.0;
  int iter_count = 1;
  if (HUTI_STOPC == HUTI_TRESID_SCALED_BYB || HUTI_STOPC == HUTI_PRESID_SCALED_BYB) rhsnorm = op.norm(ndim, B);
  auto check = [&]() {                                              // the stopping test of either half step
    if (HUTI_STOPC == HUTI_PSEUDORESIDUAL || HUTI_STOPC == HUTI_PRESID_SCALED_BYB) {
      op.matvec(X, R);
      for (int i = 0; i < ndim; ++i) R[i] = R[i] - B[i];
      op.pcondl(TRV, R);
      residual = op.norm(ndim, TRV);
      if (HUTI_STOPC == HUTI_PRESID_SCALED_BYB) residual /= rhsnorm;
    } else {
      op.pcondr(TRV, X);
      op.matvec(TRV, R);
      for (int i = 0; i < ndim; ++i) TRV[i] = R[i] - B[i];
      op.pcondl(R, TRV);
      residual = op.norm(ndim, R);
      if (HUTI_STOPC == HUTI_TRESID_SCALED_BYB) residual /= rhsnorm;
    }
  };
  auto half = [&](const double *yv, const double *av) {            // W -= alpha*av ; rotations ; D, X updates
    for (int i = 0; i < ndim; ++i) W[i] = W[i] - alpha * av[i];
    gamma = op.norm(ndim, W) / tau;
    c = 1 / std::sqrt(1 + gamma * gamma);
    tau = tau * gamma * c;
    const double f = (oldgamma * oldgamma * eta) / alpha;
    for (int i = 0; i < ndim; ++i) D[i] = yv[i] + f * D[i];
    eta = c * c * alpha;
    for (int i = 0; i < ndim; ++i) X[i] = X[i] + eta * D[i];
    oldgamma = gamma;
  };
  op.pcondr(D, X); op.matvec(D, R);
  for (int i = 0; i < ndim; ++i) D[i] = B[i] - R[i];
  op.pcondl(R, D);
  for (int i = 0; i < ndim; ++i) { Y[i] = R[i]; W[i] = R[i]; }
  op.pcondr(V, Y); op.matvec(V, D); op.pcondl(V, D);
  for (int i = 0; i < ndim; ++i) { T2V[i] = V[i]; D[i] = 0; }
  tau = op.norm(ndim, R);
  for (int i = 0; i < ndim; ++i) RTLD[i] = R[i];
  oldrho = op.dot(ndim, RTLD, R);
  if (oldrho == 0) { HUTI_INFO = 30; }                               // HUTI_TFQMR_RHO
  else for (;;) {
    alpha = oldrho / op.dot(ndim, RTLD, V);
    for (int i = 0; i < ndim; ++i) YNEW[i] = Y[i] - alpha * V[i];
    half(Y, T2V);
    check();
    if (residual < HUTI_TOLERANCE) { HUTI_INFO = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > HUTI_MAXTOLERANCE) { HUTI_INFO = HUTI_DIVERGENCE; break; }
    op.pcondr(T1V, YNEW); op.matvec(T1V, R); op.pcondl(T1V, R);
    half(YNEW, T1V);
    check();
    if (residual < HUTI_TOLERANCE) { HUTI_INFO = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > HUTI_MAXTOLERANCE) { HUTI_INFO = HUTI_DIVERGENCE; break; }
    rho = op.dot(ndim, RTLD, W);
    beta = rho / oldrho;
    for (int i = 0; i < ndim; ++i) YNEW[i] = W[i] + beta * YNEW[i];
    op.pcondr(T2V, YNEW); op.matvec(T2V, R); op.pcondl(T2V, R);
    for (int i = 0; i < ndim; ++i) V[i] = T2V[i] + beta * T1V[i] + beta * beta * V[i];
    for (int i = 0; i < ndim; ++i) Y[i] = YNEW[i];
    oldrho = rho;
    iter_count = iter_count + 1;
    if (iter_count > HUTI_MAXIT) { HUTI_INFO = HUTI_MAXITER; break; }
  }
  op.pcondr(TRV, X);
  for (int i = 0; i < ndim; ++i) X[i] = TRV[i];
  HUTI_ITERS = iter_count;
  dpar[9] = residual;
}

// fhutiter/src/huti_bicgstab_2.F90:339-578 huti_dbicgstab_2solv.  work(n,8) = RTLD,U,T1V,V,S,W,T,R.  Left-oriented in Elmer.
void huti_dbicgstab_2solv(Ops &op, int ndim, double *X, const double *B, int *ipar, double *dpar, double *work) {
  const size_t N = (size_t)ndim;
  double *RTLD = work, *U = work + N, *T1V = work + 2 * N, *V = work + 3 * N, *S = work + 4 * N, *W = work + 5 * N, *T = work + 6 * N, *R = work + 7 * N;
  double rho = 0, oldrho = 1, alpha = 0, beta = 0, omega1 = 0, omega2 = 1, tau = 0, delta = 0, myy = 0, residual = 0, rhsnorm = 1.0;
  int iter_count = 1;
  if (HUTI_STOPC == HUTI_TRESID_SCALED_BYB || HUTI_STOPC == HUTI_PRESID_SCALED_BYB) rhsnorm = op.norm(ndim, B);
  auto apply = [&](double *dst, const double *src) { op.pcondr(dst, src); op.matvec(dst, T1V); op.pcondl(dst, T1V); };   // dst = M^-1 A src
  op.pcondr(U, X); op.matvec(U, R);
  for (int i = 0; i < ndim; ++i) U[i] = B[i] - R[i];
  op.pcondl(R, U);
  for (int i = 0; i < ndim; ++i) { RTLD[i] = R[i]; U[i] = 0; }
  for (;;) {
    oldrho = -omega2 * oldrho;
    rho = op.dot(ndim, RTLD, R);
    if (rho == 0) { HUTI_INFO = 45; break; }                        // HUTI_BICGSTAB_2_RHO
    beta = (rho * alpha) / oldrho;
    oldrho = rho;
    for (int i = 0; i < ndim; ++i) U[i] = R[i] - beta * U[i];
    apply(V, U);
    alpha = oldrho / op.dot(ndim, RTLD, V);
    for (int i = 0; i < ndim; ++i) R[i] = R[i] - alpha * V[i];
    apply(S, R);
    for (int i = 0; i < ndim; ++i) X[i] = X[i] + alpha * U[i];
    rho = op.dot(ndim, RTLD, S);
    if (rho == 0) { HUTI_INFO = 45; break; }
    beta = (rho * alpha) / oldrho;
    oldrho = rho;
    for (int i = 0; i < ndim; ++i) V[i] = S[i] - beta * V[i];
    apply(W, V);
    alpha = oldrho / op.dot(ndim, RTLD, W);
    for (int i = 0; i < ndim; ++i) U[i] = R[i] - beta * U[i];
    for (int i = 0; i < ndim; ++i) R[i] = R[i] - alpha * V[i];
    for (int i = 0; i < ndim; ++i) S[i] = S[i] - alpha * W[i];
    apply(T, S);
    omega1 = op.dot(ndim, R, S);
    myy = op.dot(ndim, S, S);
    delta = op.dot(ndim, S, T);
    tau = op.dot(ndim, T, T);
    omega2 = op.dot(ndim, R, T);
    tau = tau - (delta * delta) / myy;
    omega2 = (omega2 - (delta * omega1) / myy) / tau;
    omega1 = (omega1 - delta * omega2) / myy;
    for (int i = 0; i < ndim; ++i) X[i] = X[i] + omega1 * R[i] + omega2 * S[i] + alpha * U[i];
    for (int i = 0; i < ndim; ++i) R[i] = R[i] - omega1 * S[i] - omega2 * T[i];
    if (HUTI_STOPC == HUTI_PSEUDORESIDUAL) residual = op.norm(ndim, R);
    else if (HUTI_STOPC == HUTI_PRESID_SCALED_BYB) residual = op.norm(ndim, R) / rhsnorm;
    else {
      op.pcondr(S, X); op.matvec(S, T1V);
      for (int i = 0; i < ndim; ++i) T1V[i] = T1V[i] - B[i];
      op.pcondl(S, T1V);
      residual = op.norm(ndim, S);
      if (HUTI_STOPC == HUTI_TRESID_SCALED_BYB) residual /= rhsnorm;
    }
    if (residual < HUTI_TOLERANCE) { HUTI_INFO = HUTI_CONVERGENCE; break; }
    if (residual != residual || residual > HUTI_MAXTOLERANCE) { HUTI_INFO = HUTI_DIVERGENCE; break; }
    for (int i = 0; i < ndim; ++i) U[i] = U[i] - omega1 * V[i] - omega2 * W[i];
    iter_count = iter_count + 1;
    if (iter_count > HUTI_MAXIT) { HUTI_INFO = HUTI_MAXITER; break; }
  }
  HUTI_ITERS = iter_count;
  dpar[9] = residual;
}

// fhutiter/src/huti_bicgstab.F90:279-566 huti_dbicgstabsolv.  work(n,8) = RTLD,P,T1V,V,S,T2V,T,R.
void huti_dbicgstabsolv(Ops &op, int ndim, double *X, const double *B, int *ipar, double *dpar,
                        double *work) {
  size_t N = (size_t)ndim;
  double *RTLD = work, *P = work + N, *T1V = work + 2 * N, *V = work + 3 * N, *S = work + 4 * N,
         *T2V = work + 5 * N, *T = work + 6 * N, *R = work + 7 * N;
  double rho = 0, oldrho, alpha, beta, omega, residual = 0, rhsnorm = 1.0;
  int iter_count = 1;
  if (HUTI_STOPC == HUTI_TRESID_SCALED_BYB || HUTI_STOPC == HUTI_PRESID_SCALED_BYB)
    rhsnorm = op.norm(ndim, B);
  op.matvec(X, R);
#pragma omp parallel for
  for (int i = 0; i < ndim; ++i) { R[i] = B[i] - R[i]; RTLD[i] = R[i]; }
#pragma omp parallel for
  for (int i = 0; i < ndim; ++i) { P[i] = 0; V[i] = 0; }
  oldrho = 1; omega = 1; alpha = 0;
  for (;;) {
    rho = op.dot(ndim, RTLD, R);
    if (rho == 0) { HUTI_INFO = HUTI_BICGSTAB_RHO; break; }
    beta = (rho * alpha) / (oldrho * omega);
#pragma omp parallel for
    for (int i = 0; i < ndim; ++i) P[i] = R[i] + beta * (P[i] - omega * V[i]);
    op.pcondl(V, P);
    op.pcondr(T1V, V);
    op.matvec(T1V, V);
    alpha = rho / op.dot(ndim, RTLD, V);
#pragma omp parallel for
    for (int i = 0; i < ndim; ++i) S[i] = R[i] - alpha * V[i];
    residual = op.norm(ndim, S);
    if (residual < HUTI_EPSILON) {
#pragma omp parallel for
      for (int i = 0; i < ndim; ++i) X[i] = X[i] + alpha * T1V[i];
      HUTI_INFO = HUTI_CONVERGENCE;
      break;
    }
    op.pcondl(T, S);
    op.pcondr(T2V, T);
    op.matvec(T2V, T);
    {
      double ts = op.dot(ndim, T, S);
      double tt = op.dot(ndim, T, T);
      omega = ts / tt;
    }
#pragma omp parallel for
    for (int i = 0; i < ndim; ++i) {
      X[i] = X[i] + alpha * T1V[i] + omega * T2V[i];
      R[i] = S[i] - omega * T[i];
    }
    switch (HUTI_STOPC) {
      case HUTI_TRESID_SCALED_BYB:
        op.matvec(X, T2V);
#pragma omp parallel for
        for (int i = 0; i < ndim; ++i) T1V[i] = T2V[i] - B[i];
        residual = op.norm(ndim, T1V) / rhsnorm;
        break;
      case HUTI_PSEUDORESIDUAL: residual = op.norm(ndim, R); break;
      case HUTI_PRESID_SCALED_BYB: residual = op.norm(ndim, R) / rhsnorm; break;
      case HUTI_XDIFF_NORM:
        for (int i = 0; i < ndim; ++i) T1V[i] = alpha * T1V[i] + omega * T2V[i];
        residual = op.norm(ndim, T1V);
        break;
      default:
        op.matvec(X, T2V);
#pragma omp parallel for
        for (int i = 0; i < ndim; ++i) T1V[i] = T2V[i] - B[i];
        residual = op.norm(ndim, T1V);
    }
    if (HUTI_DBUGLVL != 0 && HUTI_DBUGLVL != INT_MAX && iter_count % HUTI_DBUGLVL == 0)
      printf("%8d%11.4E\n", iter_count, residual);
    if (residual < HUTI_TOLERANCE) { HUTI_INFO = HUTI_CONVERGENCE; break; }
    if (omega == 0) { HUTI_INFO = HUTI_BICGSTAB_OMEGA; break; }
    if (residual != residual || residual > HUTI_MAXTOLERANCE) { HUTI_INFO = HUTI_DIVERGENCE; break; }
    oldrho = rho;
    iter_count = iter_count + 1;
    if (iter_count > HUTI_MAXIT) { HUTI_INFO = HUTI_MAXITER; break; }
  }
  HUTI_ITERS = iter_count;
  dpar[9] = residual;
}

// ---------------------------------------------------------------------------------------
// Tiny dense helpers standing in for the LAPACK/BLAS calls inside RealBiCGStabl
// (IterativeMethods.F90:940-1037): dgetrf/dgetrs = partial-pivot LU (row interchanges, unit
// lower), dsymv('u') = y := A*x using the upper triangle, ddot = ref_ddot.  Column-major.
struct SmallLU {
  int n; std::vector<double> a; std::vector<int> piv;
  void factor(int n_, const double *A, int lda) {
    n = n_; a.assign((size_t)n * n, 0.0); piv.assign(n, 0);
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) a[i + (size_t)j * n] = A[i + (size_t)j * lda];
    for (int j = 0; j < n; ++j) {            // dgetf2: idamax pivot, swap, scale, rank-1 update
      int p = j; double mx = std::fabs(a[j + (size_t)j * n]);
      for (int i = j + 1; i < n; ++i) if (std::fabs(a[i + (size_t)j * n]) > mx) { mx = std::fabs(a[i + (size_t)j * n]); p = i; }
      piv[j] = p;
      if (a[p + (size_t)j * n] != 0.0) {
        if (p != j) for (int k = 0; k < n; ++k) std::swap(a[j + (size_t)k * n], a[p + (size_t)k * n]);
        double r = 1.0 / a[j + (size_t)j * n];
        for (int i = j + 1; i < n; ++i) a[i + (size_t)j * n] *= r;
      }
      for (int k = j + 1; k < n; ++k)
        for (int i = j + 1; i < n; ++i) a[i + (size_t)k * n] -= a[i + (size_t)j * n] * a[j + (size_t)k * n];
    }
  }
  void solve(double *b) const {               // dgetrs 'n': dlaswp, dtrsm L unit, dtrsm U
    for (int j = 0; j < n; ++j) if (piv[j] != j) std::swap(b[j], b[piv[j]]);
    for (int j = 0; j < n; ++j) for (int i = j + 1; i < n; ++i) b[i] -= b[j] * a[i + (size_t)j * n];
    for (int j = n - 1; j >= 0; --j) {
      b[j] /= a[j + (size_t)j * n];
      for (int i = 0; i < j; ++i) b[i] -= b[j] * a[i + (size_t)j * n];
    }
  }
};
// dsymv('u', n, 1, A, lda, x, 1, 0, y, 1): reference BLAS loop order (upper triangle, column sweep).
void small_dsymv_u(int n, const double *A, int lda, const double *x, double *y) {
  for (int i = 0; i < n; ++i) y[i] = 0.0;
  for (int j = 0; j < n; ++j) {
    double temp1 = x[j], temp2 = 0.0;
    for (int i = 0; i < j; ++i) {
      y[i] = y[i] + temp1 * A[i + (size_t)j * lda];
      temp2 = temp2 + A[i + (size_t)j * lda] * x[i];
    }
    y[j] = y[j] + temp1 * A[j + (size_t)j * lda] + temp2;
  }
}

// `Linear System Robust` (IterSolve.F90:482-496; IterativeMethods.F90:607-610, 649-659, 1498-1501, 1535-1544)
struct RobustPar { bool on = false; double Tol = 0, Step = 0, MaxTol = 0; int MaxBadIter = 0, Start = 1; };
// huti_fdefs.h:132-135, 153-155: HUTI_ROBUST ipar(26), _MAXBADIT ipar(27), _START ipar(29), _TOLERANCE dpar(3), _STEPSIZE dpar(4),
// _MAXTOLERANCE dpar(5)
RobustPar robust_from(const int *ipar, const double *dpar) {
  RobustPar r;
  r.on = IPAR(26) == 1; r.MaxBadIter = IPAR(27); r.Start = IPAR(29);
  r.Tol = DPAR(3); r.Step = DPAR(4); r.MaxTol = DPAR(5);
  return r;
}

// IterativeMethods.F90:694-1168 RealBiCGStabl (no constraint matrix).
// Returns through Converged/Diverged/Halted; *rounds_out = MIN(MaxRounds, Round) as printed at 1147.
void RealBiCGStabl(Ops &op, int n, double *x, const double *b, int MaxRounds, double Tol, double MaxTol,
                   bool &Converged, bool &Diverged, bool &Halted, int OutputInterval, int l,
                   int *rounds_out, double *res_out, const RobustPar &Rb = RobustPar()) {
  const double zero = 0.0, one = 1.0, delta = 1.0e-2;
  size_t N = (size_t)n;
  // 649-659
  const bool Robust = Rb.on;
  double BestNorm = std::sqrt(DBL_MAX);
  int BadIterCount = 0, BestIter = 0;
  std::vector<double> Bestx;
  if (Robust) Bestx.assign(N, 0.0);
  (void)BestIter;
  *rounds_out = 0; *res_out = 0;
  if (l < 2) { fprintf(stderr, "RealBiCGStabl: Polynomial degree < 2\n"); Halted = true; return; }
  // 719: IF ( ALL(x == 0.0d0) ) x = b
  { bool allz = true; for (int i = 0; i < n; ++i) if (x[i] != 0.0) { allz = false; break; }
    if (allz) for (int i = 0; i < n; ++i) x[i] = b[i]; }
  const int nw = 3 + 2 * (l + 1);
  std::vector<double> workv(N * nw, 0.0), t(N, 0.0);
  const int ldr = l + 1;
  std::vector<double> rworkv((size_t)ldr * nw, 0.0);
  auto work = [&](int col) { return workv.data() + (size_t)(col - 1) * N; };          // 1-based column
  auto rwork = [&](int i, int j) -> double & { return rworkv[(i - 1) + (size_t)(j - 1) * ldr]; };
  const int rr = 1, r = rr + 1, u = r + (l + 1), xp = u + (l + 1), bp = xp + 1;
  const int z = 1, zz = z + (l + 1), y0 = zz + (l + 1), yl = y0 + 1, y = yl + 1;
  std::vector<double> tmpmtr((size_t)(l - 1) * (l - 1)), tmpvec(l - 1);
  SmallLU lu;

  op.matvec(x, work(r));
#pragma omp parallel for
  for (int i = 0; i < n; ++i) work(r)[i] = b[i] - work(r)[i];
  double bnrm = op.norm(n, b);
  double rnrm0 = op.norm(n, work(r));
  if (bnrm != bnrm || rnrm0 != rnrm0) { Diverged = true; return; }   // Fatal in the reference (769-774)
  double errorind = rnrm0 / bnrm;
  if (errorind != errorind) { Diverged = true; return; }
  *res_out = errorind;
  Converged = (errorind < Tol);
  Diverged = (errorind > MaxTol);
  if (Converged || Diverged) return;
  bool EarlyExit = false;
#pragma omp parallel for
  for (int i = 0; i < n; ++i) {
    work(rr)[i] = work(r)[i]; work(bp)[i] = work(r)[i];
    work(xp)[i] = x[i];
    x[i] = zero;
  }
  double rnrm = rnrm0, mxnrmx = rnrm0, mxnrmr = rnrm0;
  double alpha = zero, omega = one, sigma = one, rho0 = one, rho1, beta;
  int Round;
  for (Round = 1; Round <= MaxRounds; ++Round) {
    // --- The BiCG part (819-910) ---
    rho0 = -omega * rho0;
    for (int k = 1; k <= l; ++k) {
      rho1 = op.dot(n, work(rr), work(r + k - 1));
      if (rho0 == zero) { Halted = true; goto L100; }
      if (rho1 != rho1) { Diverged = true; goto L100; }            // Fatal in the reference
      beta = alpha * (rho1 / rho0);
      rho0 = rho1;
      for (int j = 0; j <= k - 1; ++j) {
        double *uj = work(u + j); const double *rj = work(r + j);
#pragma omp parallel for
        for (int i = 0; i < n; ++i) uj[i] = rj[i] - beta * uj[i];
      }
      op.pcondr(t.data(), work(u + k - 1));
      op.matvec(t.data(), work(u + k));
      sigma = op.dot(n, work(rr), work(u + k));
      if (sigma == zero) { Halted = true; goto L100; }
      if (sigma != sigma) { Diverged = true; goto L100; }          // Fatal in the reference
      alpha = rho1 / sigma;
      { const double *u0 = work(u);
#pragma omp parallel for
        for (int i = 0; i < n; ++i) x[i] = x[i] + alpha * u0[i]; }
      for (int j = 0; j <= k - 1; ++j) {
        double *rj = work(r + j); const double *uj1 = work(u + j + 1);
#pragma omp parallel for
        for (int i = 0; i < n; ++i) rj[i] = rj[i] - alpha * uj1[i];
      }
      op.pcondr(t.data(), work(r + k - 1));
      op.matvec(t.data(), work(r + k));
      rnrm = op.norm(n, work(r));
      if (rnrm != rnrm) { Diverged = true; goto L100; }            // Fatal in the reference
      mxnrmx = std::max(mxnrmx, rnrm);
      mxnrmr = std::max(mxnrmr, rnrm);
      errorind = rnrm / bnrm;
      Converged = (errorind < Tol);
      Diverged = (errorind != errorind);
      if (Converged || Diverged) { EarlyExit = true; break; }
    }
    if (EarlyExit) break;
    // --- The convex polynomial part (917-1011) ---
    for (int i = 1; i <= l + 1; ++i)
      for (int j = 1; j <= i; ++j) rwork(i, j) = op.dot(n, work(r + i - 1), work(r + j - 1));
    for (int j = 2; j <= l + 1; ++j) for (int i = 1; i <= j - 1; ++i) rwork(i, j) = rwork(j, i);
    for (int j = 0; j <= l - 1; ++j) for (int i = 1; i <= l + 1; ++i) rwork(i, zz + j) = rwork(i, z + j);
    for (int j = 1; j <= l - 1; ++j) for (int i = 1; i <= l - 1; ++i)
      tmpmtr[(i - 1) + (size_t)(j - 1) * (l - 1)] = rwork(i + 1, zz + j);
    lu.factor(l - 1, tmpmtr.data(), l - 1);
    // tilde r0 and tilde rl
    rwork(1, y0) = -one;
    for (int i = 2; i <= l; ++i) rwork(i, y0) = rwork(i, z);
    for (int i = 1; i <= l - 1; ++i) tmpvec[i - 1] = rwork(i + 1, y0);
    lu.solve(tmpvec.data());
    for (int i = 1; i <= l - 1; ++i) rwork(i + 1, y0) = tmpvec[i - 1];
    rwork(l + 1, y0) = zero;
    rwork(1, yl) = zero;
    for (int i = 1; i <= l - 1; ++i) { rwork(i + 1, yl) = rwork(i + 1, z + l); tmpvec[i - 1] = rwork(i + 1, yl); }
    lu.solve(tmpvec.data());
    for (int i = 1; i <= l - 1; ++i) rwork(i + 1, yl) = tmpvec[i - 1];
    rwork(l + 1, yl) = -one;
    // Convex combination
    double kappa0, kappal, varrho, hatgamma;
    small_dsymv_u(l + 1, &rwork(1, z), ldr, &rwork(1, y0), &rwork(1, y));
    kappa0 = ref_ddot_small(l + 1, &rwork(1, y0), &rwork(1, y));
    if (kappa0 <= 0.0) { Halted = true; goto L100; }
    kappa0 = std::sqrt(kappa0);
    small_dsymv_u(l + 1, &rwork(1, z), ldr, &rwork(1, yl), &rwork(1, y));
    kappal = ref_ddot_small(l + 1, &rwork(1, yl), &rwork(1, y));
    if (kappal <= 0.0) { Halted = true; goto L100; }
    kappal = std::sqrt(kappal);
    small_dsymv_u(l + 1, &rwork(1, z), ldr, &rwork(1, y0), &rwork(1, y));
    varrho = ref_ddot_small(l + 1, &rwork(1, yl), &rwork(1, y)) / (kappa0 * kappal);
    hatgamma = varrho / std::fabs(varrho) * std::max(std::fabs(varrho), 7e-1) * kappa0 / kappal;
    for (int i = 1; i <= l + 1; ++i) rwork(i, y0) = rwork(i, y0) - hatgamma * rwork(i, yl);
    // --- Update (1014-1033) ---
    omega = rwork(l + 1, y0);
    for (int j = 1; j <= l; ++j) {
      double g = rwork(j + 1, y0);
      double *u0 = work(u); const double *uj = work(u + j);
#pragma omp parallel for
      for (int i = 0; i < n; ++i) u0[i] = u0[i] - g * uj[i];
      const double *rj1 = work(r + j - 1);
#pragma omp parallel for
      for (int i = 0; i < n; ++i) x[i] = x[i] + g * rj1[i];
      double *r0 = work(r); const double *rj = work(r + j);
#pragma omp parallel for
      for (int i = 0; i < n; ++i) r0[i] = r0[i] - g * rj[i];
    }
    small_dsymv_u(l + 1, &rwork(1, z), ldr, &rwork(1, y0), &rwork(1, y));
    rnrm = ref_ddot_small(l + 1, &rwork(1, y0), &rwork(1, y));
    if (rnrm < 0.0) { Halted = true; goto L100; }
    rnrm = std::sqrt(rnrm);
    // --- The reliable update part (1050-1101) ---
    {
      mxnrmx = std::max(mxnrmx, rnrm);
      mxnrmr = std::max(mxnrmr, rnrm);
      bool xpdt = (rnrm < delta * rnrm0 && rnrm0 < mxnrmx);
      bool rcmp = ((rnrm < delta * mxnrmr && rnrm0 < mxnrmr) || xpdt);
      if (rcmp) {
        op.pcondr(t.data(), x);
        op.matvec(t.data(), work(r));
        mxnrmr = rnrm;
        { double *r0 = work(r); const double *bpv = work(bp);
#pragma omp parallel for
          for (int i = 0; i < n; ++i) r0[i] = bpv[i] - r0[i]; }
        if (xpdt) {
          double *xpv = work(xp), *bpv = work(bp); const double *r0 = work(r);
#pragma omp parallel for
          for (int i = 0; i < n; ++i) { xpv[i] = xpv[i] + t[i]; x[i] = zero; bpv[i] = r0[i]; }
          mxnrmx = rnrm;
        }
      }
      if (rcmp) {
        const double *xpv = work(xp);
        if (xpdt) { for (int i = 0; i < n; ++i) t[i] = xpv[i]; }
        else { for (int i = 0; i < n; ++i) t[i] = t[i] + xpv[i]; }
      } else {
        // 1095-1100: a preconditioner solve whose result is never read (kept: it is reference work)
        op.pcondr(t.data(), x);
        const double *xpv = work(xp);
        for (int i = 0; i < n; ++i) t[i] = t[i] + xpv[i];
      }
    }
    errorind = rnrm / bnrm;
    if (OutputInterval != 0 && OutputInterval != INT_MAX && Round % OutputInterval == 0)
      printf("%8d%11.4E%11.4E\n", Round, rnrm, errorind);
    if (Robust) {                                  // 1110-1126
      if (Round >= Rb.Start) {
        if (errorind < Rb.Step * BestNorm) {
          BestIter = Round;
          BestNorm = errorind;
          for (int i = 0; i < n; ++i) Bestx[i] = x[i];
          BadIterCount = 0;
        } else {
          BadIterCount = BadIterCount + 1;
        }
        if (BestNorm < Rb.Tol && (errorind > Rb.MaxTol || BadIterCount > Rb.MaxBadIter)) break;
      }
    }
    Converged = (errorind < Tol);
    Diverged = (errorind > MaxTol) || (errorind != errorind);
    if (Converged || Diverged) break;
  }
L100:
  if (Robust) {                                    // 1133-1139
    if (BestNorm < Rb.Tol) Converged = true;
    if (BestNorm < errorind) for (int i = 0; i < n; ++i) x[i] = Bestx[i];
  }
  *rounds_out = std::min(MaxRounds, Round);
  *res_out = errorind;
  // 1156-1166: x = M^-1 x + xp
#pragma omp parallel for
  for (int i = 0; i < n; ++i) t[i] = x[i];
  op.pcondr(x, t.data());
  { const double *xpv = work(xp);
#pragma omp parallel for
    for (int i = 0; i < n; ++i) x[i] = x[i] + xpv[i]; }
}

// IterativeMethods.F90:1260-1458 GCR (no constraint matrix, no pseudo-complex, no user stopc).
void GCR(Ops &op, int n, double *x, const double *b, int Rounds, double MinTolerance, double MaxTolerance,
         double &Residual, bool &Converged, bool &Diverged, int OutputInterval, int m, int MinIter,
         int *iters_out) {
  size_t N = (size_t)n;
  std::vector<double> R(N), T1(N), T2(N), S, V, trueres(N);
  if (m > 1) { S.assign(N * (m - 1), 0.0); V.assign(N * (m - 1), 0.0); }
  auto Scol = [&](int j) { return S.data() + (size_t)(j - 1) * N; };
  auto Vcol = [&](int j) { return V.data() + (size_t)(j - 1) * N; };
  double *r = R.data();
  *iters_out = 0;
  op.matvec(x, r);
  for (int i = 0; i < n; ++i) r[i] = b[i] - r[i];
  double bnorm = op.norm(n, b);
  double rnorm = op.norm(n, r);
  Residual = rnorm / bnorm;
  Converged = (Residual < MinTolerance) && (MinIter <= 0);
  Diverged = (Residual > MaxTolerance) || (Residual != Residual);
  if (Converged || Diverged) return;
  int k;
  for (k = 1; k <= Rounds; ++k) {
    int j;
    if (k % m == 0) j = m;
    else {
      j = k % m;
      if (j == 1 && k > 1) {                      // true residual when restarting (1323-1326)
        op.matvec(x, r);
        for (int i = 0; i < n; ++i) r[i] = b[i] - r[i];
      }
    }
    op.pcondr(T1.data(), r);
    op.matvec(T1.data(), T2.data());
    for (int i = 1; i <= j - 1; ++i) {            // 1338-1364
      double beta = op.dot(n, Vcol(i), T2.data());
      const double *Si = Scol(i), *Vi = Vcol(i);
      for (int q = 0; q < n; ++q) T1[q] = T1[q] - beta * Si[q];
      for (int q = 0; q < n; ++q) T2[q] = T2[q] - beta * Vi[q];
    }
    double alpha = op.norm(n, T2.data());
    { double ia = 1.0 / alpha;
      for (int q = 0; q < n; ++q) T1[q] = ia * T1[q];
      for (int q = 0; q < n; ++q) T2[q] = ia * T2[q]; }
    double beta = op.dot(n, T2.data(), r);
    for (int q = 0; q < n; ++q) x[q] = x[q] + beta * T1[q];
    for (int q = 0; q < n; ++q) r[q] = r[q] - beta * T2[q];
    if (j != m) {
      memcpy(Scol(j), T1.data(), N * sizeof(double));
      memcpy(Vcol(j), T2.data(), N * sizeof(double));
    }
    rnorm = op.norm(n, r);
    Residual = rnorm / bnorm;
    if (OutputInterval != 0 && OutputInterval != INT_MAX && k % OutputInterval == 0)
      printf("   gcr:%6d%12.4E%12.4E\n", k, Residual, beta);
    Converged = (Residual < MinTolerance) && (k >= MinIter);
    if (Converged) {                              // 1427-1431 true-residual check (informational)
      op.matvec(x, trueres.data());
      for (int i = 0; i < n; ++i) trueres[i] = b[i] - trueres[i];
      double TrueResNorm = op.norm(n, trueres.data());
      (void)TrueResNorm;
    }
    Diverged = (Residual > MaxTolerance) || (Residual != Residual);
    if (Converged || Diverged) break;
  }
  *iters_out = std::min(k, Rounds);
}

// IterativeMethods.F90:1579-1913 RealIDRS (no constraint matrix, user stopc off).
// P (n x s, column major) replaces CALL RANDOM_NUMBER(P) at 1640: the caller supplies it.
void RealIDRS(Ops &op, int n, double *x, const double *b, int MaxRounds, double Tol, double MaxTol,
              bool &Converged, bool &Diverged, int OutputInterval, int s, bool Smoothing,
              const double *Pin, int *iters_out, double *res_out, const RobustPar &Rb = RobustPar()) {
  size_t N = (size_t)n;
  // 1535-1544
  const bool Robust = Rb.on;
  double BestNorm = std::sqrt(DBL_MAX);
  int BadIterCount = 0, BestIter = 0;
  std::vector<double> Bestx;
  if (Robust) Bestx.assign(N, 0.0);
  (void)BestIter;
  std::vector<double> Pm(N * s), G(N * s, 0.0), U(N * s, 0.0), r(N), v(N), t(N);
  std::vector<double> M((size_t)s * s, 0.0), f(s), mu(s), alpha(s), beta(s), gamma(s);
  std::vector<double> r_s, x_s;
  auto P = [&](int j) { return Pm.data() + (size_t)(j - 1) * N; };
  auto Gc = [&](int j) { return G.data() + (size_t)(j - 1) * N; };
  auto Uc = [&](int j) { return U.data() + (size_t)(j - 1) * N; };
  auto Mm = [&](int i, int j) -> double & { return M[(i - 1) + (size_t)(j - 1) * s]; };
  double om, tr, tr_s, tt, nr, nt, rho, kappa, theta, normb, normr, errorind;
  int iter = 0, ii = 0, jj = 0;
  *iters_out = 0;
  normb = op.norm(n, b);
  op.matvec(x, t.data());
  for (int i = 0; i < n; ++i) r[i] = b[i] - t[i];
  normr = op.norm(n, r.data());
  errorind = normr / normb;
  *res_out = errorind;
  Converged = (errorind < Tol);
  Diverged = (errorind > MaxTol) || (errorind != errorind);
  if (Converged || Diverged) return;
  if (Smoothing) { x_s.assign(x, x + n); r_s = r; }
  memcpy(Pm.data(), Pin, N * s * sizeof(double));
  for (int j = 1; j <= s; ++j) {                  // 1655-1661 Gram-Schmidt on P
    for (int k = 1; k <= j - 1; ++k) {
      alpha[k - 1] = op.dot(n, P(k), P(j));
      double a = alpha[k - 1]; double *pj = P(j); const double *pk = P(k);
      for (int i = 0; i < n; ++i) pj[i] = pj[i] - a * pk[i];
    }
    double nrm = op.norm(n, P(j)); double *pj = P(j);
    for (int i = 0; i < n; ++i) pj[i] = pj[i] / nrm;
  }
  kappa = 0.7;
  om = 1.0;
  while (!Converged && !Diverged) {
    for (int k = 1; k <= s; ++k) f[k - 1] = op.dot(n, P(k), r.data());
    for (int k = 1; k <= s; ++k) {
      ii = ii + 1;
      for (int i = 0; i < n; ++i) v[i] = r[i];
      if (jj > 0) {
        for (int i = k; i <= s; ++i) {            // 1696-1703
          gamma[i - 1] = f[i - 1];
          for (int j = k; j <= i - 1; ++j) gamma[i - 1] = gamma[i - 1] - Mm(i, j) * gamma[j - 1];
          gamma[i - 1] = gamma[i - 1] / Mm(i, i);
          double g = gamma[i - 1]; const double *Gi = Gc(i);
          for (int q = 0; q < n; ++q) v[q] = v[q] - g * Gi[q];
        }
        op.pcondr(t.data(), v.data());
        for (int q = 0; q < n; ++q) t[q] = om * t[q];
        for (int i = k; i <= s; ++i) {
          double g = gamma[i - 1]; const double *Ui = Uc(i);
          for (int q = 0; q < n; ++q) t[q] = t[q] + g * Ui[q];
        }
        memcpy(Uc(k), t.data(), N * sizeof(double));
      } else {
        op.pcondr(Uc(k), v.data());
      }
      op.matvec(Uc(k), Gc(k));
      for (int i = 1; i <= s; ++i) mu[i - 1] = op.dot(n, P(i), Gc(k));
      for (int i = 1; i <= k - 1; ++i) {          // 1727-1736
        alpha[i - 1] = mu[i - 1];
        for (int j = 1; j <= i - 1; ++j) alpha[i - 1] = alpha[i - 1] - Mm(i, j) * alpha[j - 1];
        alpha[i - 1] = alpha[i - 1] / Mm(i, i);
        double a = alpha[i - 1]; double *Gk = Gc(k), *Uk = Uc(k); const double *Gi = Gc(i), *Ui = Uc(i);
        for (int q = 0; q < n; ++q) Gk[q] = Gk[q] - Gi[q] * a;
        for (int q = 0; q < n; ++q) Uk[q] = Uk[q] - Ui[q] * a;
        for (int q = k; q <= s; ++q) mu[q - 1] = mu[q - 1] - Mm(q, i) * a;
      }
      for (int q = k; q <= s; ++q) Mm(q, k) = mu[q - 1];
      if (std::fabs(Mm(k, k)) <= DBL_MIN) { Diverged = true; break; }   // TINY(tol)
      beta[k - 1] = f[k - 1] / Mm(k, k);
      { double bk = beta[k - 1]; const double *Gk = Gc(k), *Uk = Uc(k);
        for (int q = 0; q < n; ++q) r[q] = r[q] - bk * Gk[q];
        for (int q = 0; q < n; ++q) x[q] = x[q] + bk * Uk[q]; }
      if (k < s) for (int q = k + 1; q <= s; ++q) f[q - 1] = f[q - 1] - beta[k - 1] * Mm(q, k);
      if (Smoothing) {
        for (int q = 0; q < n; ++q) t[q] = r_s[q] - r[q];
        tr_s = op.dot(n, t.data(), r_s.data());
        tt = op.dot(n, t.data(), t.data());
        theta = tr_s / tt;
        for (int q = 0; q < n; ++q) r_s[q] = r_s[q] - theta * t[q];
        for (int q = 0; q < n; ++q) x_s[q] = x_s[q] - theta * (x_s[q] - x[q]);
      }
      iter = iter + 1;
      normr = Smoothing ? op.norm(n, r_s.data()) : op.norm(n, r.data());
      errorind = normr / normb;
      if (OutputInterval != 0 && OutputInterval != INT_MAX && iter % OutputInterval == 0)
        printf("%8d%11.4E\n", iter, errorind);
      Converged = (errorind < Tol);
      Diverged = (errorind > MaxTol) || (errorind != errorind);
      if (Converged || Diverged) break;
      if (iter == MaxRounds) break;
    }
    if (Converged || Diverged) break;
    if (iter == MaxRounds) break;
    jj = jj + 1;
    op.pcondr(v.data(), r.data());
    op.matvec(v.data(), t.data());
    nr = op.norm(n, r.data());
    nt = op.norm(n, t.data());
    tr = op.dot(n, t.data(), r.data());
    rho = std::fabs(tr / (nt * nr));
    om = tr / (nt * nt);
    if (rho < kappa) om = om * kappa / rho;
    if (std::fabs(om) <= DBL_EPSILON) { Diverged = true; break; }
    for (int q = 0; q < n; ++q) r[q] = r[q] - om * t[q];
    for (int q = 0; q < n; ++q) x[q] = x[q] + om * v[q];
    if (Smoothing) {
      for (int q = 0; q < n; ++q) t[q] = r_s[q] - r[q];
      tr_s = op.dot(n, t.data(), r_s.data());
      tt = op.dot(n, t.data(), t.data());
      theta = tr_s / tt;
      for (int q = 0; q < n; ++q) r_s[q] = r_s[q] - theta * t[q];
      for (int q = 0; q < n; ++q) x_s[q] = x_s[q] - theta * (x_s[q] - x[q]);
    }
    iter = iter + 1;
    normr = Smoothing ? op.norm(n, r_s.data()) : op.norm(n, r.data());
    errorind = normr / normb;
    if (OutputInterval != 0 && OutputInterval != INT_MAX && iter % OutputInterval == 0)
      printf("%8d%11.4E\n", iter, errorind);
    if (Robust) {                                  // 1862-1884
      if (errorind < Rb.Step * BestNorm) {
        BestIter = iter;
        BestNorm = errorind;
        const double *src = Smoothing ? x_s.data() : x;
        for (int q = 0; q < n; ++q) Bestx[q] = src[q];
        BadIterCount = 0;
      } else {
        BadIterCount = BadIterCount + 1;
      }
      if (BestNorm < Rb.Tol && (errorind > Rb.MaxTol || BadIterCount > Rb.MaxBadIter)) break;
    }
    Converged = (errorind < Tol);
    Diverged = (errorind > MaxTol) || (errorind != errorind);
    if (iter == MaxRounds) break;
  }
  if (Smoothing) for (int q = 0; q < n; ++q) x[q] = x_s[q];
  if (Robust) {                                    // 1892-1898
    if (BestNorm < Rb.Tol) Converged = true;
    if (BestNorm < errorind) for (int q = 0; q < n; ++q) x[q] = Bestx[q];
  }
  *iters_out = iter;
  *res_out = errorind;
}

}  // namespace

// =======================================================================================
// C entry points (ctypes).  method: 1 cg, 2 bicgstab, 3 bicgstabl, 4 gcr, 5 idrs.
// precond: 0 none, 1 diagonal, 2 ilu0.
extern "C" {

void orc_set_dot_order(int mode) { g_dot_order = mode; }
void orc_set_device_blocks(int blocks) { g_dev_blocks = blocks; }
void orc_set_threads(int nthreads) {
#ifdef _OPENMP
  omp_set_num_threads(nthreads > 0 ? nthreads : 1);
#else
  (void)nthreads;
#endif
}
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

double orc_ddot(int n, const double *x, const double *y) { return ref_ddot(n, x, y); }
double orc_dnrm2(int n, const double *x) { return ref_dnrm2(n, x); }

void orc_crs_matvec(int n, const int *rows, const int *cols, const double *vals, int ndeg,
                    const double *u, double *v) {
  Matrix A{n, rows, cols, nullptr, vals, nullptr, ndeg, 0, 0, 0, 0, 0};
  crs_matvec(A, u, v);
}
void orc_crs_diag_precond(int n, const int *rows, const int *cols, const int *diag, const double *vals,
                          double *u, const double *v) {
  Matrix A{n, rows, cols, diag, vals, nullptr, 1, 1, 0, 0, 0, 0};
  crs_diag_precond(A, u, v);
}
int orc_crs_ilu0(int n, const int *rows, const int *cols, const int *diag, const double *vals,
                 double *iluvals) {
  return crs_ilu0(n, rows, cols, diag, vals, iluvals);
}
void orc_set_cholesky(int flag) { g_cholesky = flag; }
long orc_crs_ilut(int n, const int *rows, const int *cols, const double *vals, double tol, long cap, int *ilurows, int *ilucols, int *iludiag,
                  double *iluvals) {
  return crs_ilut(n, rows, cols, vals, tol, cap, ilurows, ilucols, iludiag, iluvals);
}
int orc_crs_ichol_factor(int n, const int *rows, const int *cols, const int *diag, const double *vals, const int *ilurows, const int *ilucols,
                         const int *iludiag, double *iluvals) {
  return crs_ichol_factor(n, rows, cols, diag, vals, ilurows, ilucols, iludiag, iluvals);
}
long orc_crs_ilu1_pattern(int n, const int *rows, const int *cols, const int *diag, int *ilurows, int *ilucols, int *iludiag) {
  return crs_ilu1_pattern(n, rows, cols, diag, ilurows, ilucols, iludiag);
}
int orc_crs_ilun_factor(int n, const int *rows, const int *cols, const double *vals, const int *ilurows, const int *ilucols,
                        const int *iludiag, double *iluvals) {
  return crs_ilun_factor(n, rows, cols, vals, ilurows, ilucols, iludiag, iluvals);
}
// same as orc_itersolve with the ILU(n) factor given on its own pattern
int orc_itersolve(int n, const int *rows, const int *cols, const int *diag, const double *vals, int ndeg,
                  const double *iluvals, const double *b, double *x, int *ipar, double *dpar,
                  int method, int precond, const double *P, long *counts);
static const int *g_ilu_rows = nullptr, *g_ilu_cols = nullptr, *g_ilu_diag = nullptr;
int orc_itersolve_ilun(int n, const int *rows, const int *cols, const int *diag, const double *vals, int ndeg,
                       const int *ilurows, const int *ilucols, const int *iludiag, const double *iluvals, const double *b, double *x,
                       int *ipar, double *dpar, int method, const double *P, long *counts) {
  g_ilu_rows = ilurows; g_ilu_cols = ilucols; g_ilu_diag = iludiag;
  int rc = orc_itersolve(n, rows, cols, diag, vals, ndeg, iluvals, b, x, ipar, dpar, method, 2, P, counts);
  g_ilu_rows = g_ilu_cols = g_ilu_diag = nullptr;
  return rc;
}
void orc_crs_lu_precond(int n, const int *rows, const int *cols, const int *diag, const double *iluvals,
                        double *u, const double *v) {
  Matrix A{n, rows, cols, diag, nullptr, iluvals, 1, 2, 0, 0, 0, 0};
  crs_lu_precond(A, u, v);
}

// The part of IterSolver (IterSolve.F90:159-1047) that surrounds the method call, for a real CRS
// matrix with right-oriented preconditioning: x = 1e-8 rule (470-471), DBUGLVL 0 -> HUGE for the
// internal methods (913), dispatch (865-897), and the INFO mapping of the internal methods
// (IterativeMethods.F90:676-684, 1253-1255, 1567-1569).  ipar/dpar must be filled by the caller
// exactly as IterSolver fills them.  iluvals: precomputed ILU0 factor (orc_crs_ilu0) or NULL.
// P: n x s shadow space for idrs.  counts[4] out: matvec, pcond, dot, norm calls.
// ipar(31) is set to the iteration count for every method (the reference leaves it 0 for the three
// IterativeMethods routines; they only print it).
int orc_itersolve(int n, const int *rows, const int *cols, const int *diag, const double *vals, int ndeg,
                  const double *iluvals, const double *b, double *x, int *ipar, double *dpar,
                  int method, int precond, const double *P, long *counts) {
  Matrix A{n, rows, cols, diag, vals, iluvals, ndeg, precond, 0, 0, 0, 0};
  A.ILURows = g_ilu_rows; A.ILUCols = g_ilu_cols; A.ILUDiag = g_ilu_diag;
  Ops op{&A};
  HUTI_NDIM = n;
  if (method == 2 || method == 3 || method == 9) {
    bool allz = true;
    for (int i = 0; i < n; ++i) if (x[i] != 0.0) { allz = false; break; }
    if (allz) for (int i = 0; i < n; ++i) x[i] = 1.0e-8;
  }
  if (method >= 3 && HUTI_DBUGLVL == 0) HUTI_DBUGLVL = INT_MAX;
  double res = 0;
  if (method == 1) {
    std::vector<double> work((size_t)n * 4, 0.0);
    huti_dcgsolv(op, n, x, b, ipar, dpar, work.data());
  } else if (method == 2) {
    std::vector<double> work((size_t)n * 8, 0.0);
    huti_dbicgstabsolv(op, n, x, b, ipar, dpar, work.data());
  } else if (method == 3) {
    bool Converged = false, Diverged = false, Halted = false; int rounds = 0;
    RealBiCGStabl(op, n, x, b, HUTI_MAXIT, HUTI_TOLERANCE, HUTI_MAXTOLERANCE, Converged, Diverged, Halted,
                  HUTI_DBUGLVL, HUTI_BICGSTABL_L, &rounds, &res, robust_from(ipar, dpar));
    if (Converged) HUTI_INFO = HUTI_CONVERGENCE;
    else if (Diverged) HUTI_INFO = HUTI_DIVERGENCE;
    else if (Halted) HUTI_INFO = HUTI_HALTED;
    else HUTI_INFO = HUTI_MAXITER;
    HUTI_ITERS = rounds; dpar[9] = res;
  } else if (method == 4) {
    bool Converged = false, Diverged = false; int iters = 0;
    GCR(op, n, x, b, HUTI_MAXIT, HUTI_TOLERANCE, HUTI_MAXTOLERANCE, res, Converged, Diverged,
        HUTI_DBUGLVL, HUTI_GCR_RESTART, HUTI_MINIT, &iters);
    if (Converged) HUTI_INFO = HUTI_CONVERGENCE;
    if (Diverged) HUTI_INFO = HUTI_DIVERGENCE;
    if (!Converged && !Diverged) HUTI_INFO = HUTI_MAXITER;
    HUTI_ITERS = iters; dpar[9] = res;
  } else if (method == 5) {
    bool Converged = false, Diverged = false; int iters = 0;
    RealIDRS(op, n, x, b, HUTI_MAXIT, HUTI_TOLERANCE, HUTI_MAXTOLERANCE, Converged, Diverged,
             HUTI_DBUGLVL, HUTI_IDRS_S, HUTI_SMOOTHING == 1, P, &iters, &res, robust_from(ipar, dpar));
    if (Converged) HUTI_INFO = HUTI_CONVERGENCE;
    if (Diverged) HUTI_INFO = HUTI_DIVERGENCE;
    if (!Converged && !Diverged) HUTI_INFO = HUTI_MAXITER;
    HUTI_ITERS = iters; dpar[9] = res;
  } else if (method == 12) {
    bool Converged = false, Diverged = false; int iters = 0;
    SGS(op, n, x, b, HUTI_MAXIT, HUTI_TOLERANCE, HUTI_MAXTOLERANCE, res, Converged, Diverged, DPAR(3), &iters);
    if (Converged) HUTI_INFO = HUTI_CONVERGENCE;
    if (Diverged) HUTI_INFO = HUTI_DIVERGENCE;
    if (!Converged && !Diverged) HUTI_INFO = HUTI_MAXITER;
    HUTI_ITERS = iters; dpar[9] = res;
  } else if (method == 10 || method == 11) {
    bool Converged = false, Diverged = false; int iters = 0;
    StationaryIteration(op, n, x, b, HUTI_MAXIT, HUTI_TOLERANCE, HUTI_MAXTOLERANCE, res, Converged, Diverged, method == 11, &iters);
    if (Converged) HUTI_INFO = HUTI_CONVERGENCE;
    if (Diverged) HUTI_INFO = HUTI_DIVERGENCE;
    if (!Converged && !Diverged) HUTI_INFO = HUTI_MAXITER;
    HUTI_ITERS = iters; dpar[9] = res;
  } else if (method == 9) {
    op.left = true;
    std::vector<double> work((size_t)n * 8, 0.0);
    huti_dbicgstab_2solv(op, n, x, b, ipar, dpar, work.data());
  } else if (method == 7) {
    std::vector<double> work((size_t)n * 7, 0.0);
    huti_dcgssolv(op, n, x, b, ipar, dpar, work.data());
  } else if (method == 8) {
    op.left = true;
    std::vector<double> work((size_t)n * 10, 0.0);
    huti_dtfqmrsolv(op, n, x, b, ipar, dpar, work.data());
  } else if (method == 6) {
    op.left = true;
    std::vector<double> work((size_t)n * (7 + HUTI_GMRES_RESTART), 0.0);
    huti_dgmressolv(op, n, x, b, ipar, dpar, work.data());
  } else {
    return -1;
  }
  if (counts) { counts[0] = A.n_matvec; counts[1] = A.n_pcond; counts[2] = A.n_dot; counts[3] = A.n_norm; }
  return 0;
}

// SolverUtils.F90:12976-13213 ScaleLinearSystemDiagonal (real, serial, no PrecValues/Mass/Damp):
// Diag = 1/sqrt(|a_ii|) (row abs-sum if a_ii is tiny), Values *= D_i D_j, b *= D, b /= ||b||,
// Diag *= ||b||, x /= Diag.  Returns bnorm (A%RhsScaling).
double orc_scale_system(int n, const int *rows, const int *cols, const int *diagp, double *vals,
                        double *b, double *x, double *Diag) {
  const double tiny = DBL_MIN;
  for (int i = 0; i < n; ++i) { Diag[i] = 0.0; int j = diagp[i]; if (j > 0) Diag[i] = vals[j - 1]; }
  bool anytiny = false;
  for (int i = 0; i < n; ++i) if (std::fabs(Diag[i]) <= tiny) anytiny = true;
  if (anytiny) {
    for (int i = 0; i < n; ++i) {
      if (std::fabs(Diag[i]) <= tiny) {
        double s = 0; for (int j = rows[i] - 1; j < rows[i + 1] - 1; ++j) s += std::fabs(vals[j]);
        Diag[i] = s;
      }
    }
  }
  for (int i = 0; i < n; ++i) Diag[i] = (std::fabs(Diag[i]) > tiny) ? 1.0 / std::sqrt(std::fabs(Diag[i])) : 1.0;
#pragma omp parallel for
  for (int i = 0; i < n; ++i)
    for (int j = rows[i] - 1; j < rows[i + 1] - 1; ++j) vals[j] = vals[j] * (Diag[i] * Diag[cols[j] - 1]);
  for (int i = 0; i < n; ++i) b[i] = b[i] * Diag[i];
  double bnorm = 0; for (int i = 0; i < n; ++i) bnorm += b[i] * b[i];
  bnorm = std::sqrt(bnorm);
  bool DoRhs = true;
  if (bnorm < std::sqrt(tiny)) { DoRhs = false; bnorm = 1.0; }
  if (DoRhs) { for (int i = 0; i < n; ++i) { Diag[i] = Diag[i] * bnorm; b[i] = b[i] / bnorm; } }
  for (int i = 0; i < n; ++i) x[i] = x[i] / Diag[i];
  return bnorm;
}

// SolverUtils.F90:13515-13643 BackScaleLinearSystemDiagonal: x *= Diag; Diag /= bnorm;
// b = b/Diag*bnorm; Values /= D_i D_j.
void orc_backscale_system(int n, const int *rows, const int *cols, double *vals, double *b, double *x,
                          double *Diag, double bnorm) {
  for (int i = 0; i < n; ++i) x[i] = x[i] * Diag[i];
  for (int i = 0; i < n; ++i) Diag[i] = Diag[i] / bnorm;
  for (int i = 0; i < n; ++i) b[i] = b[i] / Diag[i] * bnorm;
#pragma omp parallel for
  for (int i = 0; i < n; ++i)
    for (int j = rows[i] - 1; j < rows[i + 1] - 1; ++j) vals[j] = vals[j] / (Diag[i] * Diag[cols[j] - 1]);
}

}  // extern "C"

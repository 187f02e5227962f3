// HOST TOOL, not on the solve path: synthetic-input generator for tests and bench.py (the workloads
// BASELINE.json names are "synthetic ElmerGrid-generated meshes").  It produces what Elmer would hand
// to IterSolver: structured hex8 meshes numbered exactly as
// ElmerGrid numbers a one-subcell .grd with "Numbering = Horizontal" (SURVEY.md Appendix A,
// checked against the reference's own ElmerGrid output in tests/test_meshgen_vs_elmergrid.py),
// the CRS structure Elmer's CreateMatrix produces for nodal dofs without bandwidth optimisation
// (rows = ndof*(node-1)+c, sorted columns, ElementUtils.F90:1659-1690), hex8 element matrices with
// the default 2x2x2 Gauss rule (fem/src/elements.def:607-641; Poisson form as
// fem/tests/PoissonThreaded/Poisson.F90:165-203), and Dirichlet rows as EnforceDirichletConditions
// leaves them under the default scaling (SolverUtils.F90:8702-8745: row zeroed, diagonal kept at
// |a_kk|, b = a_kk * value; optional symmetric elimination as CRS_ElimSymmDirichlet).
//
// All index arrays are 1-based like Elmer's.

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// Reference hex8 nodes (elements.def 808): (-1,-1,-1),(1,-1,-1),(1,1,-1),(-1,1,-1),(-1,-1,1),(1,-1,1),(1,1,1),(-1,1,1)
const double HU[8] = {-1, 1, 1, -1, -1, 1, 1, -1};
const double HV[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
const double HW[8] = {-1, -1, -1, -1, 1, 1, 1, 1};

struct GP { double N[8], dN[8][3], w; };   // dN wrt physical coordinates, w = weight*detJ

// Shape functions and physical gradients at the 8 Gauss points of one hex8 element.
void hex8_gauss(const double X[8][3], GP gp[8]) {
  const double g = 1.0 / std::sqrt(3.0);
  int q = 0;
  for (int c = 0; c < 2; ++c) for (int b = 0; b < 2; ++b) for (int a = 0; a < 2; ++a, ++q) {
    double u = a ? g : -g, v = b ? g : -g, w = c ? g : -g;
    double dNu[8][3];
    for (int p = 0; p < 8; ++p) {
      gp[q].N[p] = 0.125 * (1 + HU[p] * u) * (1 + HV[p] * v) * (1 + HW[p] * w);
      dNu[p][0] = 0.125 * HU[p] * (1 + HV[p] * v) * (1 + HW[p] * w);
      dNu[p][1] = 0.125 * (1 + HU[p] * u) * HV[p] * (1 + HW[p] * w);
      dNu[p][2] = 0.125 * (1 + HU[p] * u) * (1 + HV[p] * v) * HW[p];
    }
    double J[3][3] = {{0}};
    for (int p = 0; p < 8; ++p) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J[i][j] += dNu[p][i] * X[p][j];
    double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                 J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    double Ji[3][3];
    Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
    Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    Ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det;
    Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    for (int p = 0; p < 8; ++p) for (int i = 0; i < 3; ++i) {
      double s = 0; for (int j = 0; j < 3; ++j) s += Ji[i][j] * dNu[p][j];
      gp[q].dN[p][i] = s;
    }
    gp[q].w = std::fabs(det);
  }
}

// kind 0: Poisson/heat  K_pq = sum w grad(N_p).grad(N_q), F_p = sum w f N_p            (ndof 1)
// kind 1: isotropic linear elasticity (E, nu), body force (fx,fy,fz)                      (ndof 3)
// kind 2: Picard-linearised convection-diffusion-pressure block system (u,v,w,p)          (ndof 4)
//         momentum rows: nu*(grad,grad) + (a.grad N_q) N_p (+ SUPG-like streamline term), pressure
//         coupling -p div(v) / q div(u) and a PSPG-like pressure Laplacian; a = lid-driven-like field.
struct Params { int kind; double c[8]; };

void element_matrix(const Params &P, const double X[8][3], int ndof, double *K /*(8*ndof)^2 row-major*/, double *F) {
  GP gp[8];
  hex8_gauss(X, gp);
  const int m = 8 * ndof;
  for (int i = 0; i < m * m; ++i) K[i] = 0;
  for (int i = 0; i < m; ++i) F[i] = 0;
  if (P.kind == 0) {
    for (int q = 0; q < 8; ++q) {
      for (int p = 0; p < 8; ++p) {
        for (int r = 0; r < 8; ++r)
          K[p * 8 + r] += gp[q].w * (gp[q].dN[p][0] * gp[q].dN[r][0] + gp[q].dN[p][1] * gp[q].dN[r][1] + gp[q].dN[p][2] * gp[q].dN[r][2]);
        F[p] += gp[q].w * P.c[0] * gp[q].N[p];
      }
    }
  } else if (P.kind == 1) {
    const double E = P.c[0], nu = P.c[1];
    const double lam = E * nu / ((1 + nu) * (1 - 2 * nu)), mu = E / (2 * (1 + nu));
    for (int q = 0; q < 8; ++q) {
      for (int p = 0; p < 8; ++p) {
        for (int r = 0; r < 8; ++r) {
          const double *a = gp[q].dN[p], *b = gp[q].dN[r];
          double ab = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
          for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            double v = lam * a[i] * b[j] + mu * a[j] * b[i] + (i == j ? mu * ab : 0.0);
            K[(p * 3 + i) * m + (r * 3 + j)] += gp[q].w * v;
          }
        }
        for (int i = 0; i < 3; ++i) F[p * 3 + i] += gp[q].w * P.c[2 + i] * gp[q].N[p];
      }
    }
  } else {
    const double visc = P.c[0], tau = P.c[1];
    for (int q = 0; q < 8; ++q) {
      double xq[3] = {0, 0, 0};
      for (int p = 0; p < 8; ++p) for (int i = 0; i < 3; ++i) xq[i] += gp[q].N[p] * X[p][i];
      // smooth recirculating field resembling a lid-driven cavity (divergence free in x-z)
      const double pi = 3.14159265358979323846;
      double a[3] = {std::sin(pi * xq[0]) * std::sin(pi * xq[0]) * std::sin(2 * pi * xq[2]) * xq[2],
                     0.1 * std::sin(2 * pi * xq[0]) * std::sin(2 * pi * xq[1]),
                     -std::sin(2 * pi * xq[0]) * std::sin(pi * xq[2]) * std::sin(pi * xq[2]) * xq[2]};
      for (int p = 0; p < 8; ++p) {
        const double *gpN = gp[q].dN[p];
        double adp = a[0] * gpN[0] + a[1] * gpN[1] + a[2] * gpN[2];
        for (int r = 0; r < 8; ++r) {
          const double *gr = gp[q].dN[r];
          double adr = a[0] * gr[0] + a[1] * gr[1] + a[2] * gr[2];
          double diff = visc * (gpN[0] * gr[0] + gpN[1] * gr[1] + gpN[2] * gr[2]);
          double conv = adr * gp[q].N[p] + tau * adr * adp;
          for (int i = 0; i < 3; ++i) {
            K[(p * 4 + i) * m + (r * 4 + i)] += gp[q].w * (diff + conv);
            K[(p * 4 + i) * m + (r * 4 + 3)] += gp[q].w * (-gpN[i] * gp[q].N[r] + tau * adp * gr[i]);   // pressure gradient
            K[(p * 4 + 3) * m + (r * 4 + i)] += gp[q].w * (gp[q].N[p] * gr[i] + tau * gpN[i] * adr);    // continuity
          }
          K[(p * 4 + 3) * m + (r * 4 + 3)] += gp[q].w * tau * (gpN[0] * gr[0] + gpN[1] * gr[1] + gpN[2] * gr[2]);
        }
        F[p * 4 + 2] += gp[q].w * P.c[2] * gp[q].N[p];
      }
    }
  }
}

}  // namespace

extern "C" {

// Nodes x-fastest: id(a,b,c) = 1 + a + (ex+1)*b + (ex+1)*(ey+1)*c; elements x-fastest with
// connectivity bottom face counter-clockwise then top face (ElmerGrid probe: "1 2 43 42 1682 1683 1724 1723").
void fem_grid_hex8(int ex, int ey, int ez, double lx, double ly, double lz, double *xyz, int *elems) {
  const int nx = ex + 1, ny = ey + 1, nz = ez + 1;
#pragma omp parallel for
  for (int c = 0; c < nz; ++c) for (int b = 0; b < ny; ++b) for (int a = 0; a < nx; ++a) {
    size_t id = a + (size_t)nx * b + (size_t)nx * ny * c;
    xyz[3 * id + 0] = lx * a / ex; xyz[3 * id + 1] = ly * b / ey; xyz[3 * id + 2] = lz * c / ez;
  }
#pragma omp parallel for
  for (int c = 0; c < ez; ++c) for (int b = 0; b < ey; ++b) for (int a = 0; a < ex; ++a) {
    size_t e = a + (size_t)ex * b + (size_t)ex * ey * c;
    int n1 = 1 + a + nx * b + nx * ny * c;
    int *E = elems + 8 * e;
    E[0] = n1; E[1] = n1 + 1; E[2] = n1 + 1 + nx; E[3] = n1 + nx;
    E[4] = n1 + nx * ny; E[5] = n1 + 1 + nx * ny; E[6] = n1 + 1 + nx + nx * ny; E[7] = n1 + nx + nx * ny;
  }
}

// node -> element adjacency (CSR, 0-based element ids). ptr has nn+1 entries.
static void node_elems(int nn, long ne, const int *elems, int nen, std::vector<long> &ptr, std::vector<int> &adj) {
  ptr.assign((size_t)nn + 1, 0);
  for (long e = 0; e < ne; ++e) for (int k = 0; k < nen; ++k) ptr[elems[e * nen + k]]++;
  for (int i = 0; i < nn; ++i) ptr[i + 1] += ptr[i];
  adj.resize(ptr[nn]);
  std::vector<long> pos(ptr.begin(), ptr.end() - 1);
  for (long e = 0; e < ne; ++e) for (int k = 0; k < nen; ++k) adj[pos[elems[e * nen + k] - 1]++] = (int)e;
}

// Pass 1: rows[0..n] (1-based pointers) for ndof interleaved dofs per node. Returns nnz.
long fem_crs_count(int nn, long ne, const int *elems, int nen, int ndof, int *rows) {
  std::vector<long> ptr; std::vector<int> adj;
  node_elems(nn, ne, elems, nen, ptr, adj);
  std::vector<int> cnt(nn);
#pragma omp parallel
  {
    std::vector<int> nb;
#pragma omp for
    for (int i = 0; i < nn; ++i) {
      nb.clear();
      for (long q = ptr[i]; q < ptr[i + 1]; ++q) for (int k = 0; k < nen; ++k) nb.push_back(elems[(long)adj[q] * nen + k]);
      std::sort(nb.begin(), nb.end());
      cnt[i] = (int)(std::unique(nb.begin(), nb.end()) - nb.begin());
    }
  }
  long nnz = 0;
  rows[0] = 1;
  for (int i = 0; i < nn; ++i) for (int c = 0; c < ndof; ++c) {
    nnz += (long)cnt[i] * ndof;
    if (nnz + 1 > 2147483647L) return -nnz;
    rows[(size_t)i * ndof + c + 1] = (int)(nnz + 1);
  }
  return nnz;
}

// Pass 2: cols (sorted ascending per row) and diag (1-based position of the diagonal entry).
void fem_crs_fill(int nn, long ne, const int *elems, int nen, int ndof, const int *rows, int *cols, int *diag) {
  std::vector<long> ptr; std::vector<int> adj;
  node_elems(nn, ne, elems, nen, ptr, adj);
#pragma omp parallel
  {
    std::vector<int> nb;
#pragma omp for
    for (int i = 0; i < nn; ++i) {
      nb.clear();
      for (long q = ptr[i]; q < ptr[i + 1]; ++q) for (int k = 0; k < nen; ++k) nb.push_back(elems[(long)adj[q] * nen + k]);
      std::sort(nb.begin(), nb.end());
      nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
      for (int c = 0; c < ndof; ++c) {
        size_t r = (size_t)i * ndof + c;
        int p = rows[r] - 1;
        for (size_t t = 0; t < nb.size(); ++t) for (int d = 0; d < ndof; ++d) {
          int col = ndof * (nb[t] - 1) + d + 1;
          cols[p] = col;
          if (col == (int)r + 1) diag[r] = p + 1;
          ++p;
        }
      }
    }
  }
}

// Row-wise assembly (no atomics): each dof row visits its adjacent elements.  uniform != 0 reuses
// the element matrix of element 0 (valid when all elements are congruent, kinds 0 and 1 only).
void fem_assemble(int kind, const double *par, int nn, long ne, const int *elems, const double *xyz, int ndof,
                  const int *rows, const int *cols, double *vals, double *rhs, int uniform) {
  const int nen = 8, m = nen * ndof;
  Params P; P.kind = kind; for (int i = 0; i < 8; ++i) P.c[i] = par[i];
  std::vector<long> ptr; std::vector<int> adj;
  node_elems(nn, ne, elems, nen, ptr, adj);
  std::vector<double> K0((size_t)m * m), F0(m);
  auto coords = [&](long e, double X[8][3]) {
    for (int k = 0; k < 8; ++k) for (int d = 0; d < 3; ++d) X[k][d] = xyz[3 * (size_t)(elems[e * 8 + k] - 1) + d];
  };
  if (uniform) { double X[8][3]; coords(0, X); element_matrix(P, X, ndof, K0.data(), F0.data()); }
#pragma omp parallel
  {
    std::vector<double> K((size_t)m * m), F(m);
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < nn; ++i) {
      for (int c = 0; c < ndof; ++c) {
        size_t r = (size_t)i * ndof + c;
        for (int p = rows[r] - 1; p < rows[r + 1] - 1; ++p) vals[p] = 0.0;
        rhs[r] = 0.0;
      }
      for (long q = ptr[i]; q < ptr[i + 1]; ++q) {
        long e = adj[q];
        const double *Ke = K0.data(), *Fe = F0.data();
        if (!uniform) { double X[8][3]; coords(e, X); element_matrix(P, X, ndof, K.data(), F.data()); Ke = K.data(); Fe = F.data(); }
        int lp = 0; for (int k = 0; k < 8; ++k) if (elems[e * 8 + k] == i + 1) lp = k;
        for (int c = 0; c < ndof; ++c) {
          size_t r = (size_t)i * ndof + c;
          const int *cb = cols + rows[r] - 1; int len = rows[r + 1] - rows[r];
          rhs[r] += Fe[lp * ndof + c];
          for (int k = 0; k < 8; ++k) {
            int col0 = ndof * (elems[e * 8 + k] - 1) + 1;
            int pos = (int)(std::lower_bound(cb, cb + len, col0) - cb);
            for (int d = 0; d < ndof; ++d) vals[rows[r] - 1 + pos + d] += Ke[(lp * ndof + c) * m + (k * ndof + d)];
          }
        }
      }
    }
  }
}

// Dirichlet rows (EnforceDirichletConditions under default scaling): optional symmetric elimination
// b_i -= a_ik*val for the other rows (CRS_ElimSymmDirichlet), then row k zeroed except the diagonal,
// which keeps s = |a_kk| (s = 1/DiagScaling^2), and b_k = s*val.
void fem_dirichlet(int n, const int *rows, const int *cols, const int *diag, double *vals, double *b,
                   int ndir, const int *dofs, const double *dvals, int symmetric) {
  std::vector<char> isd(n, 0); std::vector<double> dv(n, 0.0);
  for (int t = 0; t < ndir; ++t) { isd[dofs[t] - 1] = 1; dv[dofs[t] - 1] = dvals[t]; }
  if (symmetric) {
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
      if (isd[i]) continue;
      for (int p = rows[i] - 1; p < rows[i + 1] - 1; ++p) {
        int c = cols[p] - 1;
        if (isd[c]) { b[i] -= vals[p] * dv[c]; vals[p] = 0.0; }
      }
    }
  }
#pragma omp parallel for
  for (int i = 0; i < n; ++i) {
    if (!isd[i]) continue;
    double s = std::fabs(vals[diag[i] - 1]);
    if (s <= 2.2250738585072014e-308) s = 1.0;
    for (int p = rows[i] - 1; p < rows[i + 1] - 1; ++p) vals[p] = 0.0;
    vals[diag[i] - 1] = s;
    b[i] = s * dv[i];
  }
}

// Linear System Scaling (default TRUE, SolverUtils.F90:14492-14498): what ScaleLinearSystemDiagonal
// (SolverUtils.F90:12976-13213) leaves in A, b before IterSolver is called, for a real system without
// constraints: D = 1/sqrt(|a_ii|), A <- D A D, b <- D b / ||D b||.  Returns ||D b||; D (times that norm)
// is kept so that the caller can scale x back.  The drop-in never runs this: Elmer does it on the host.
double fem_scale_system(int n, const int *rows, const int *cols, const int *diag, double *vals, double *b, double *D) {
  const double tiny = 2.2250738585072014e-308;
#pragma omp parallel for
  for (int i = 0; i < n; ++i) {
    double d = std::fabs(vals[diag[i] - 1]);
    if (d <= tiny) { d = 0; for (int p = rows[i] - 1; p < rows[i + 1] - 1; ++p) d += std::fabs(vals[p]); }
    D[i] = d > tiny ? 1.0 / std::sqrt(d) : 1.0;
  }
#pragma omp parallel for
  for (int i = 0; i < n; ++i)
    for (int p = rows[i] - 1; p < rows[i + 1] - 1; ++p) vals[p] = vals[p] * (D[i] * D[cols[p] - 1]);
  double s = 0;
  for (int i = 0; i < n; ++i) { b[i] = b[i] * D[i]; s += b[i] * b[i]; }
  double bnorm = std::sqrt(s);
  if (bnorm < std::sqrt(tiny)) return 1.0;
  for (int i = 0; i < n; ++i) { D[i] = D[i] * bnorm; b[i] = b[i] / bnorm; }
  return bnorm;
}

}  // extern "C"

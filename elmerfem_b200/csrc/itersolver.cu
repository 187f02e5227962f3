// Host-side mirror of IterSolver(A,x,b,Solver), fem/src/IterSolve.F90:159-1047: keyword parsing,
// ipar/dpar filling, preconditioner recompute policy and dispatch -- everything IterSolver does
// around the HUTI call, for the keyword combinations this library implements.  Anything else is
// DECLINED (the caller then runs Elmer's own path); declining is not a fallback inside the library.
#include "common.cuh"
#include "krylov.h"
#include "../../include/elmer_b200.h"
#include <map>
#include <string>
#include <sstream>
#include <algorithm>
#include <cctype>
#include <climits>

namespace b200 {

// SIF keyword names are case-insensitive and whitespace-insensitive between words (Lists.F90 lowercases
// and single-spaces them on read).
static std::string norm_key(const std::string &s) {
  std::string o; bool sp = false;
  for (char ch : s) {
    if (isspace((unsigned char)ch)) { sp = !o.empty(); continue; }
    if (sp) { o += ' '; sp = false; }
    o += (char)tolower((unsigned char)ch);
  }
  return o;
}
static std::string strip_type(std::string v) {
  // "Real 1e-8", "Integer 500", "Logical True", "String ilu0", quoted strings
  std::string n = norm_key(v);
  for (const char *t : {"real ", "integer ", "logical ", "string "}) if (n.rfind(t, 0) == 0) { n = n.substr(strlen(t)); break; }
  if (n.size() >= 2 && n.front() == '"' && n.back() == '"') n = n.substr(1, n.size() - 2);
  return n;
}

struct Sif {
  std::map<std::string, std::string> kv;
  explicit Sif(const char *text) {
    std::istringstream in(text ? text : "");
    std::string line;
    while (std::getline(in, line)) {
      size_t c = line.find('!'); if (c != std::string::npos) line = line.substr(0, c);
      size_t e = line.find('=');
      if (e == std::string::npos) continue;
      std::string k = norm_key(line.substr(0, e));
      size_t dc = k.find("::"); if (dc != std::string::npos) k = norm_key(k.substr(dc + 2));   // "Solver 1 :: key"
      size_t pr = k.find('('); if (pr != std::string::npos) k = norm_key(k.substr(0, pr));
      kv[k] = strip_type(line.substr(e + 1));
    }
  }
  bool has(const char *k) const { return kv.count(norm_key(k)) > 0; }
  std::string str(const char *k, const char *d, bool *found = nullptr) const {
    auto it = kv.find(norm_key(k)); if (found) *found = it != kv.end();
    return it == kv.end() ? std::string(d) : it->second;
  }
  bool logical(const char *k, bool d = false, bool *found = nullptr) const {
    auto it = kv.find(norm_key(k)); if (found) *found = it != kv.end();
    if (it == kv.end()) return d;
    return it->second == "true" || it->second == "1" || it->second == ".true.";
  }
  int integer(const char *k, int d, bool *found = nullptr) const {
    auto it = kv.find(norm_key(k)); if (found) *found = it != kv.end();
    return it == kv.end() ? d : (int)strtol(it->second.c_str(), nullptr, 10);
  }
  double real(const char *k, double d, bool *found = nullptr) const {
    auto it = kv.find(norm_key(k)); if (found) *found = it != kv.end();
    if (it == kv.end()) return d;
    std::string v = it->second; std::replace(v.begin(), v.end(), 'd', 'e');   // Fortran 1.0d-8
    return strtod(v.c_str(), nullptr);
  }
};

struct Declined { std::string why; };

}  // namespace b200


namespace b200 {

// What IterSolver decides from the keywords alone (IterSolve.F90:250-577): method, preconditioner, ipar/dpar.  Pure host logic --
// b200_itersolver_plan exposes it without a device.  Throws Declined for keyword combinations the library does not implement,
// BEFORE anything is touched.
struct SolvePlan {
  int method = 0, pc = 0;
  int ipar[50] = {0}; double dpar[10] = {0};
  int ilu_order = -1;       // >= 0 when an ILU(n) / BILU preconditioner is selected
  int bilu_blocks = 0;      // > 1: ILU(0) of the block-diagonal part with that many blocks
  bool cholesky = false;    // Linear System Symmetric ILU
  bool ilut = false; double ilut_tol = 0.0;   // Linear System Preconditioning = ILUT, Linear System ILUT Tolerance
};

static void plan_from_sif(const Sif &P, int n, int ndeg, SolvePlan &pl) {
  bool found;
  // ---- method (IterSolve.F90:250-315)
  std::string m = P.str("Linear System Iterative Method", "bicgstab", &found);
  int &method = pl.method;
  if (m == "cg") method = B200_METHOD_CG;
  else if (m == "bicgstab") method = B200_METHOD_BICGSTAB;
  else if (m == "bicgstabl") method = B200_METHOD_BICGSTABL;
  else if (m == "gcr") method = B200_METHOD_GCR;
  else if (m == "idrs") method = B200_METHOD_IDRS;
  else if (m == "gmres") method = B200_METHOD_GMRES;
  else if (m == "cgs") method = B200_METHOD_CGS;
  else if (m == "tfqmr") method = B200_METHOD_TFQMR;
  else if (m == "bicgstab2") method = B200_METHOD_BICGSTAB2;
  else if (m == "jacobi") method = B200_METHOD_JACOBI;
  else if (m == "richardson") method = B200_METHOD_RICHARDSON;
  else if (m == "sgs") method = B200_METHOD_SGS;
  else method = B200_METHOD_BICGSTAB;                                  // CASE DEFAULT (313-314): unknown names run BiCGStab
  if (P.logical("Linear System Complex") || P.logical("Linear System Pseudo Complex"))
    throw Declined{"complex / pseudo-complex systems"};
  const bool internal = (method >= B200_METHOD_BICGSTABL && method <= B200_METHOD_IDRS) || method == B200_METHOD_JACOBI || method == B200_METHOD_RICHARDSON || method == B200_METHOD_SGS;
  // ---- work sizes and method parameters (327-392)
  pl.ipar[3] = internal ? 1 : (method == B200_METHOD_CG ? 4 : 8);
  if (method == B200_METHOD_CGS) pl.ipar[3] = 7;                         // HUTI_CGS_WORKSIZE
  if (method == B200_METHOD_TFQMR) pl.ipar[3] = 10;                      // HUTI_TFQMR_WORKSIZE
  if (method == B200_METHOD_BICGSTAB2) pl.ipar[3] = 8;                   // HUTI_BICGSTAB_2_WORKSIZE
  if (method == B200_METHOD_GMRES) {                                  // 346-350
    int r = P.integer("Linear System GMRES Restart", 10, &found);
    B200_REQUIRE(r >= 1, "'Linear System GMRES Restart' < 1");
    pl.ipar[14] = r; pl.ipar[3] = 7 + r;
  }
  const int maxit = P.integer("Linear System Max Iterations", 0, &found);
  B200_REQUIRE(found && maxit >= 1, "'Linear System Max Iterations' missing or < 1");
  if (method == B200_METHOD_GCR) {
    int r = P.integer("Linear System GCR Restart", 0, &found);
    if (!found) r = std::min(maxit, 200);                             // 369-376
    pl.ipar[16] = r;
  }
  if (method == B200_METHOD_BICGSTABL) {
    int l = P.integer("BiCGstabl polynomial degree", 2, &found);
    B200_REQUIRE(!found || l >= 2, "'BiCGstabl polynomial degree' < 2");
    pl.ipar[15] = l;
  }
  if (method == B200_METHOD_SGS) {                                    // 354-359: the default is the single-precision literal 1.8
    double om = P.real("SGS Overrelaxation Factor", 0.0, &found);
    pl.dpar[2] = found ? om : (double)1.8f;
  }
  if (method == B200_METHOD_IDRS) {
    int s = P.integer("IDRS parameter", 4, &found);
    B200_REQUIRE(!found || s >= 1, "'IDRS parameter' < 1");
    pl.ipar[17] = s;
  }
  // ---- stopping criterion (397-433)
  if (P.logical("Linear System Componentwise Backward Error") || P.logical("Linear System Normwise Backward Error"))
    throw Declined{"backward-error stopping criteria"};
  pl.ipar[11] = 1;                                                        // HUTI_TRESID_SCALED_BYB
  pl.ipar[2] = n;                                                       // HUTI_NDIM
  pl.ipar[4] = P.integer("Linear System Residual Output", 1, &found);    // HUTI_DBUGLVL (436-438)
  pl.ipar[9] = maxit;
  pl.ipar[10] = P.integer("Linear System Min Iterations", 0);
  pl.ipar[13] = 1;                                                        // HUTI_USERSUPPLIEDX (473)
  pl.dpar[0] = P.real("Linear System Convergence Tolerance", 0.0);
  pl.dpar[1] = P.real("Linear System Divergence Limit", 1.0e20, &found);
  if (P.logical("Linear System Robust")) {                             // 482-496, defaults as there; after the SGS factor, which
    pl.ipar[25] = 1;                                                      // shares dpar(3) with the robust tolerance
    pl.dpar[2] = P.real("Linear System Robust Tolerance", 0.0, &found);
    if (!found) pl.dpar[2] = pow(pl.dpar[0], (double)(2.0f / 3.0f));         // HUTI_TOLERANCE**(2.0/3.0): default-real exponent
    pl.dpar[4] = P.real("Linear System Robust Limit", 0.0, &found);
    if (!found) pl.dpar[4] = sqrt(pl.dpar[0]);
    pl.dpar[3] = P.real("Linear System Robust Margin", 0.0, &found);
    if (!found) pl.dpar[3] = 1.1;
    pl.ipar[26] = P.integer("Linear System Robust Max Iterations", 0, &found);
    if (!found) pl.ipar[26] = maxit / 2;
    pl.ipar[28] = P.integer("Linear System Robust Start Iteration", 0, &found);
    if (!found) pl.ipar[28] = 1;
  }
  pl.ipar[27] = P.logical("IDRS Smoothing") ? 1 : 0;
  // ---- preconditioner (506-577)
  // GMRES is left-preconditioned by IterSolver itself (509-525) and so is run_gmres; for CG/BiCGStab the keyword is declined
  if (!internal && method != B200_METHOD_GMRES && method != B200_METHOD_TFQMR && method != B200_METHOD_BICGSTAB2 && P.logical("Linear System Left Preconditioning"))
    throw Declined{"left-oriented preconditioning"};
  std::string pcs = P.str("Linear System Preconditioning", "none");
  int &pc = pl.pc;
  pl.cholesky = P.logical("Linear System Symmetric ILU");             // 526: A % Cholesky
  if (pcs == "none") pc = B200_PRECOND_NONE;
  else if (pcs == "diagonal") pc = B200_PRECOND_DIAGONAL;
  else if (pcs == "ilut") {                                           // 535-538: PRECOND_ILUT, tolerance 0 when the keyword is absent
    pl.ilut = true; pl.ilut_tol = P.real("Linear System ILUT Tolerance", 0.0);
    pl.ilu_order = 0; pl.bilu_blocks = 0;
    pc = B200_PRECOND_ILU0;
  }
  else if (pcs.rfind("ilu", 0) == 0) {
    int ilun; bool got; double o = P.real("Linear System ILU Order", 0.0, &got);
    if (got) ilun = (int)lround(o);
    else ilun = pcs.size() >= 4 ? pcs[3] - '0' : -1;                  // 541-546
    if (ilun < 0 || ilun > 9) ilun = 0;
    pl.ilu_order = ilun; pl.bilu_blocks = 0;
    pc = B200_PRECOND_ILU0;
  } else if (pcs.rfind("bilu", 0) == 0) {
    // IterSolve.F90:549-558, 745-765: ILU(n) of the block-diagonal part, Blocks = Solver % Variable % Dofs (here: Matrix_t % ndeg,
    // or the shim's "B200 Variable Dofs").  Order 0 only: for n > 0 the reference's RE-factorisation leaves stale entries of the
    // scattered row behind (only pattern positions of S are cleared, CRSMatrix.F90:3643-3649) and is not reproducible as a preconditioner.
    int ilun = pcs.size() >= 5 ? pcs[4] - '0' : 0;
    if (ilun < 0 || ilun > 9) ilun = 0;
    if (ilun != 0) throw Declined{"BILU order > 0"};
    int blocks = P.integer("B200 Variable Dofs", ndeg);
    if (blocks <= 1) blocks = 0;
    pl.ilu_order = 0; pl.bilu_blocks = blocks;
    pc = B200_PRECOND_ILU0;
  } else if (pcs == "multigrid" || pcs.rfind("vanka", 0) == 0 || pcs == "slave" || pcs == "circuit")
    throw Declined{"preconditioner '" + pcs + "'"};
  else { fprintf(stderr, "[elmer_b200] IterSolve: Unknown preconditioner type, feature disabled.\n"); pc = B200_PRECOND_NONE; }
  if (P.real("Linear System ILU Factor", 0.0) > 2.220446049250313e-16) throw Declined{"'Linear System ILU Factor'"};
  if (P.logical("Edge Basis")) throw Declined{"'Edge Basis' preconditioner matrix"};
}

}  // namespace b200

using namespace b200;

// Host-only: the decisions b200_itersolver would take for these keywords on a matrix of n rows with ndeg dofs per node.
extern "C" int b200_itersolver_plan(const char *sif, const int *n, const int *ndeg, int *method, int *precond, int *ilu_order,
                                    int *bilu_blocks, int *ipar, double *dpar) {
  try {
    B200_REQUIRE(n && ndeg && method && precond && ipar && dpar, "b200_itersolver_plan: null argument");
    Sif P(sif);
    SolvePlan pl;
    plan_from_sif(P, *n, *ndeg, pl);
    *method = pl.method; *precond = pl.pc;
    if (ilu_order) *ilu_order = pl.ilu_order;
    if (bilu_blocks) *bilu_blocks = pl.bilu_blocks;
    std::copy(pl.ipar, pl.ipar + 50, ipar); std::copy(pl.dpar, pl.dpar + 10, dpar);
    return 0;
  } catch (const Declined &d) {
    set_last_error("declined: " + d.why);
    return B200_DECLINED;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    return 1;
  }
}

extern "C" int b200_itersolver(void **handle, const double *b, double *x, const char *sif, int *solve_count, int *info_out) {
  try {
    B200_REQUIRE(handle && *handle, "null handle");
    Handle &h = *static_cast<Handle *>(*handle);
    Sif P(sif);
    SolvePlan pl;
    plan_from_sif(P, h.n, h.ndeg, pl);
    int *ipar = pl.ipar; double *dpar = pl.dpar;
    int method = pl.method, pc = pl.pc;
    if (pl.ilu_order >= 0 && (pl.ilu_order != h.ilu_order || pl.bilu_blocks != h.bilu_blocks || pl.cholesky != h.cholesky || pl.ilut != h.ilut ||
                              (pl.ilut && pl.ilut_tol != h.ilut_tol))) {
      B200_CUDA(cudaSetDevice(h.device));
      h.ilu_order = pl.ilu_order; h.bilu_blocks = pl.bilu_blocks; h.cholesky = pl.cholesky; h.ilut = pl.ilut; h.ilut_tol = pl.ilut_tol;
      ilu_invalidate(h);
    }
    // ---- recompute policy (579-587): factorise when no factor exists or Refactorize and SolveCount mod n == 0
    int sc = solve_count ? *solve_count : 0;
    if (pc == B200_PRECOND_ILU0) {
      B200_CUDA(cudaSetDevice(h.device));
      bool refactor = false;
      if (!P.logical("No Precondition Recompute")) {
        int n = P.integer("Linear System Precondition Recompute", 1); if (n <= 0) n = 1;
        bool Refactorize = P.logical("Linear System Refactorize", true);
        refactor = !h.ilu_exists || (Refactorize && sc % n == 0);
      }
      if (refactor || !h.ilu_exists) ilu0_factor(h);
      h.ilu_valid = true;                                               // a stale factor is reused on purpose
    }
    if (solve_count) *solve_count = sc + 1;                             // A % SolveCount (787)
    int rc = b200_solve(handle, b, x, ipar, dpar, &method, &pc, nullptr);
    if (info_out) { info_out[0] = ipar[29]; info_out[1] = ipar[30]; }
    return rc;
  } catch (const Declined &d) {
    set_last_error("declined: " + d.why);
    return B200_DECLINED;
  } catch (const std::exception &e) {
    set_last_error(e.what());
    fprintf(stderr, "[elmer_b200] %s\n", e.what());
    if (info_out) { info_out[0] = B200_INFO_HALTED; info_out[1] = 0; }
    return 1;
  }
}

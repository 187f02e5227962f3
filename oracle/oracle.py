"""TEST INFRASTRUCTURE -- ctypes front-end of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
Nothing under elmerfem_b200/ does.

The Python side only marshals arrays; every numerical loop is in elmer_oracle.cpp,
each citing the reference file:line it restates.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

METHODS = {"cg": 1, "bicgstab": 2, "bicgstabl": 3, "gcr": 4, "idrs": 5, "gmres": 6, "cgs": 7, "tfqmr": 8, "bicgstab2": 9, "jacobi": 10, "richardson": 11, "sgs": 12}
PRECONDS = {"none": 0, "diagonal": 1, "ilu0": 2, "ilu": 2, "ilu1": 2, "ilu2": 2, "ilu3": 2}

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/elmergrid/src") and not os.path.exists(os.path.join(_HERE, "_ref", "ElmerGrid")):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.orc_ddot.restype = C.c_double
    L.orc_ddot.argtypes = [C.c_int, _dp, _dp]
    L.orc_dnrm2.restype = C.c_double
    L.orc_dnrm2.argtypes = [C.c_int, _dp]
    L.orc_crs_matvec.argtypes = [C.c_int, _ip, _ip, _dp, C.c_int, _dp, _dp]
    L.orc_crs_diag_precond.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp, _dp]
    L.orc_crs_ilu0.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp]
    L.orc_crs_lu_precond.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp, _dp]
    L.orc_crs_ilu1_pattern.restype = C.c_long
    L.orc_crs_ilu1_pattern.argtypes = [C.c_int, _ip, _ip, _ip, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_crs_ilun_factor.argtypes = [C.c_int, _ip, _ip, _dp, _ip, _ip, _ip, _dp]
    L.orc_itersolve_ilun.argtypes = [C.c_int, _ip, _ip, _ip, _dp, C.c_int, _ip, _ip, _ip, _dp, _dp, _dp, _ip, _dp, C.c_int,
                                     C.c_void_p, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")]
    L.orc_itersolve.argtypes = [C.c_int, _ip, _ip, _ip, _dp, C.c_int, C.c_void_p, _dp, _dp, _ip, _dp,
                                C.c_int, C.c_int, C.c_void_p, _lp]
    L.orc_scale_system.restype = C.c_double
    L.orc_scale_system.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp, _dp, _dp]
    L.orc_backscale_system.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, C.c_double]
    L.orc_make_list_matrix.restype = C.c_long
    L.orc_make_list_matrix.argtypes = [C.c_int, _ip, _ip, _ip, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_optimize_bandwidth.argtypes = [C.c_int, _ip, _ip, C.c_int, _ip, _ip, C.c_int, C.c_int]
    L.orc_initialize_matrix.argtypes = [C.c_int, _ip, _ip, C.c_int, C.c_void_p, C.c_void_p, _ip, _ip, _ip]
    L.orc_set_cholesky.argtypes = [C.c_int]
    L.orc_crs_ilut.restype = C.c_long
    L.orc_crs_ilut.argtypes = [C.c_int, _ip, _ip, _dp, C.c_double, C.c_long, _ip, _ip, _ip, _dp]
    L.orc_crs_ichol_factor.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _ip, _ip, _ip, _dp]
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_max_threads.restype = C.c_int
    _LIB = L
    return L


def set_threads(n):
    lib().orc_set_threads(int(n))


def set_dot_order(mode):
    """0 = the reference's ddot/dnrm2 (default); 1, 2 = the same dot products summed in the orders other BLAS
    builds use (sensitivity study only, see elmer_oracle.cpp); 3 = the summation order of the device reductions
    (grid-stride thread partials, xor-shuffle trees, block partials: elmerfem_b200/csrc/common.cuh grid_reduce), under
    which a single-rank device solve must agree with the oracle bit for bit."""
    lib().orc_set_dot_order(int(mode))


_CHOLESKY = False


def set_cholesky(flag):
    """A % Cholesky ('Linear System Symmetric ILU', IterSolve.F90:526): ilu0 / ilun / lu_precond / itersolve take the incomplete
    Cholesky branches of CRS_IncompleteLU (CRSMatrix.F90:3539-3602) and CRS_LUSolve (4618-4638)."""
    global _CHOLESKY
    _CHOLESKY = bool(flag)
    lib().orc_set_cholesky(int(bool(flag)))


def set_device_blocks(blocks=148 * 8):
    """Grid size of the device reductions emulated by dot order 3 (B200_BLAS_BLOCKS / B200_SPMV_BLOCKS, default 148 * 8)."""
    lib().orc_set_device_blocks(int(blocks))


def max_threads():
    return lib().orc_max_threads()


from elmerfem_b200.synth import (CRS, grid_hex8, crs_structure, assemble, dirichlet, boundary_nodes,   # noqa: E402,F401
                                 heat_cube, elasticity_beam, cavity_flow)


# ------------------------------------------------------------------ kernels
def matvec(A, u):
    v = np.empty(A.n)
    lib().orc_crs_matvec(A.n, A.rows, A.cols, A.vals, A.ndeg, np.ascontiguousarray(u, dtype=np.float64), v)
    return v


def diag_precond(A, v):
    u = np.empty(A.n)
    lib().orc_crs_diag_precond(A.n, A.rows, A.cols, A.diag, A.vals, u, np.ascontiguousarray(v, dtype=np.float64))
    return u


def ilu0(A):
    ilu = np.zeros(A.nnz)
    if _CHOLESKY:
        lib().orc_crs_ichol_factor(A.n, A.rows, A.cols, A.diag, A.vals, A.rows, A.cols, A.diag, ilu)
    else:
        lib().orc_crs_ilu0(A.n, A.rows, A.cols, A.diag, A.vals, ilu)
    return ilu


def ilun(A, order):
    """CRS_IncompleteLU(A, order) for order >= 1 (CRSMatrix.F90:3445-3795): `order` rounds of InitializeILU1 on the
    pattern, then the row-by-row factorisation.  Returns a CRS holding ILURows/ILUCols/ILUDiag/ILUValues."""
    rows, cols, diag = A.rows, A.cols, A.diag
    for _ in range(order):
        nz = lib().orc_crs_ilu1_pattern(A.n, rows, cols, diag, None, None, None)
        r2 = np.zeros(A.n + 1, dtype=np.int32); c2 = np.zeros(nz, dtype=np.int32); d2 = np.zeros(A.n, dtype=np.int32)
        lib().orc_crs_ilu1_pattern(A.n, rows, cols, diag, r2.ctypes.data_as(C.c_void_p), c2.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p))
        rows, cols, diag = r2, c2, d2
    vals = np.zeros(cols.size)
    if _CHOLESKY:
        lib().orc_crs_ichol_factor(A.n, A.rows, A.cols, A.diag, A.vals, rows, cols, diag, vals)
    else:
        lib().orc_crs_ilun_factor(A.n, A.rows, A.cols, A.vals, rows, cols, diag, vals)
    return CRS(rows, cols, diag, vals, A.ndeg)


def ilut(A, tol):
    """CRS_ILUT(A, TOL) (CRSMatrix.F90:4144-4340): returns a CRS holding ILURows/ILUCols/ILUDiag/ILUValues (pattern decided by the values)."""
    cap = max(4 * A.nnz, 1024)
    while True:
        r = np.zeros(A.n + 1, dtype=np.int32); c = np.zeros(cap, dtype=np.int32); d = np.zeros(A.n, dtype=np.int32); v = np.zeros(cap)
        nz = lib().orc_crs_ilut(A.n, A.rows, A.cols, A.vals, float(tol), cap, r, c, d, v)
        if nz >= 0:
            return CRS(r, c[:nz].copy(), d, v[:nz].copy(), A.ndeg)
        cap *= 4


def lu_precond(A, ilu, v):
    """ilu: ILUValues on A's pattern (ILU0), or the CRS returned by ilun()."""
    if isinstance(ilu, CRS):
        A, ilu = ilu, ilu.vals
    u = np.empty(A.n)
    lib().orc_crs_lu_precond(A.n, A.rows, A.cols, A.diag, ilu, u, np.ascontiguousarray(v, dtype=np.float64))
    return u


def ddot(x, y):
    return lib().orc_ddot(x.size, np.ascontiguousarray(x), np.ascontiguousarray(y))


def dnrm2(x):
    return lib().orc_dnrm2(x.size, np.ascontiguousarray(x))


# ------------------------------------------------------------------ IterSolver
def fill_ipar_dpar(n, method, tol=1e-8, maxit=1000, minit=0, maxtol=1e20, residual_output=0,
                   bicgstabl_l=2, gcr_restart=None, idrs_s=4, smoothing=False, stopc=1, gmres_restart=10, sgs_omega=None,
                   robust=False, robust_tol=None, robust_limit=None, robust_margin=None, robust_max_bad=None, robust_start=None):
    """ipar/dpar exactly as IterSolver fills them (IterSolve.F90:245-503), HUTI slots per
    fhutiter/src/huti_fdefs.h:101-155."""
    ipar = np.zeros(50, dtype=np.int32)
    dpar = np.zeros(10, dtype=np.float64)
    wrk = {"cg": 4, "bicgstab": 8, "cgs": 7, "tfqmr": 10, "bicgstab2": 8}.get(method, 1)
    ipar[3 - 1] = n
    ipar[4 - 1] = wrk
    ipar[5 - 1] = residual_output
    ipar[10 - 1] = maxit
    ipar[11 - 1] = minit
    ipar[12 - 1] = stopc              # HUTI_TRESID_SCALED_BYB (IterSolve.F90:430)
    ipar[14 - 1] = 1                  # HUTI_USERSUPPLIEDX (473)
    if method == "gmres":             # IterSolve.F90:346-350
        ipar[15 - 1] = gmres_restart
        ipar[4 - 1] = 7 + gmres_restart
    if method == "bicgstabl":
        ipar[16 - 1] = max(2, bicgstabl_l)
    if method == "gcr":
        ipar[17 - 1] = gcr_restart if gcr_restart is not None else min(maxit, 200)
    if method == "idrs":
        ipar[18 - 1] = idrs_s
    ipar[28 - 1] = 1 if smoothing else 0
    dpar[1 - 1] = tol
    dpar[3 - 1] = float(np.float32(1.8)) if sgs_omega is None else sgs_omega     # HUTI_SGSPARAM; the default is the REAL literal 1.8 (IterSolve.F90:358)
    dpar[2 - 1] = maxtol
    if robust:                        # IterSolve.F90:482-496 (after the SGS factor: dpar(3) is shared)
        ipar[26 - 1] = 1
        dpar[3 - 1] = tol ** float(np.float32(2.0) / np.float32(3.0)) if robust_tol is None else robust_tol   # **(2.0/3.0), default real
        dpar[5 - 1] = np.sqrt(tol) if robust_limit is None else robust_limit
        dpar[4 - 1] = 1.1 if robust_margin is None else robust_margin
        ipar[27 - 1] = maxit // 2 if robust_max_bad is None else robust_max_bad
        ipar[29 - 1] = 1 if robust_start is None else robust_start
    return ipar, dpar


def shadow_space(n, s, seed=314159265):
    """Stand-in for CALL RANDOM_NUMBER(P) (IterativeMethods.F90:1640): uniform [0,1), column major
    (n x s).  gfortran's generator cannot be reproduced here; the same P goes to oracle and GPU."""
    rng = np.random.RandomState(seed % (2 ** 32))
    return np.asfortranarray(rng.random_sample((n, s)))


def itersolve(A, b, x0=None, method="bicgstab", precond="none", ilu=None, P=None, **kw):
    """IterSolver on the oracle.  Returns dict(x, info, iters, residual, counts)."""
    n = A.n
    x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    b = np.ascontiguousarray(b, dtype=np.float64)
    ipar, dpar = fill_ipar_dpar(n, method, **kw)
    pc = PRECONDS[precond]
    order = {"ilu1": 1, "ilu2": 2, "ilu3": 3}.get(precond, 0)
    if order and ilu is None:
        ilu = ilun(A, order)
    if pc == 2 and ilu is None:
        ilu = ilu0(A)
    ilu_p = ilu.ctypes.data_as(C.c_void_p) if (pc == 2 and not isinstance(ilu, CRS)) else None
    Pp = None
    if method == "idrs":
        s = int(ipar[18 - 1])
        if P is None:
            P = shadow_space(n, s)
        P = np.asfortranarray(P, dtype=np.float64)
        Pp = P.ctypes.data_as(C.c_void_p)
    counts = np.zeros(4, dtype=np.int64)
    if isinstance(ilu, CRS):      # ILU(n) factor on its own pattern
        rc = lib().orc_itersolve_ilun(n, A.rows, A.cols, A.diag, A.vals, A.ndeg, ilu.rows, ilu.cols, ilu.diag, ilu.vals, b, x, ipar, dpar,
                                      METHODS[method], Pp, counts)
    else:
        rc = lib().orc_itersolve(n, A.rows, A.cols, A.diag, A.vals, A.ndeg, ilu_p, b, x, ipar, dpar,
                                 METHODS[method], pc, Pp, counts)
    assert rc == 0
    return dict(x=x, info=int(ipar[30 - 1]), iters=int(ipar[31 - 1]), residual=float(dpar[9]),
                counts=dict(matvec=int(counts[0]), pcond=int(counts[1]), dot=int(counts[2]), norm=int(counts[3])),
                ipar=ipar, dpar=dpar)


def scale_system(A, b, x):
    """In place ScaleLinearSystemDiagonal; returns (Diag, bnorm)."""
    D = np.zeros(A.n)
    bn = lib().orc_scale_system(A.n, A.rows, A.cols, A.diag, A.vals, b, x, D)
    return D, bn


def backscale_system(A, b, x, D, bnorm):
    lib().orc_backscale_system(A.n, A.rows, A.cols, A.vals, b, x, D, bnorm)


def solve_linear_system(A, b, x0=None, scaling=True, **kw):
    """SolveLinearSystem's iterative branch (SolverUtils.F90:14748-14751, 14869, 14925-14927):
    default diagonal scaling, IterSolver, back-scaling.  A and b are left as on entry."""
    A = A.copy()
    b = np.array(b, dtype=np.float64, copy=True)
    x = np.zeros(A.n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    if scaling:
        D, bn = scale_system(A, b, x)
    r = itersolve(A, b, x, **kw)
    x = r["x"]
    if scaling:
        backscale_system(A, b, x, D, bn)
    r["x"] = x
    r["norm"] = float(np.sqrt(np.sum(x * x) / A.n))   # ComputeNorm, SolverUtils.F90:10290
    return r


def make_list_matrix(elem_ptr, elem_nodes, reorder, k):
    """MakeListMatrix, plain nodal branch (ElementUtils.F90:881-891): 1-based CRS of the list matrix."""
    ep = np.ascontiguousarray(elem_ptr, dtype=np.int32); en = np.ascontiguousarray(elem_nodes, dtype=np.int32)
    ro = np.ascontiguousarray(reorder, dtype=np.int32)
    nnz = lib().orc_make_list_matrix(ep.size - 1, ep, en, ro, k, None, None)
    rows = np.zeros(k + 1, dtype=np.int32); cols = np.zeros(max(nnz, 1), dtype=np.int32)
    lib().orc_make_list_matrix(ep.size - 1, ep, en, ro, k, rows.ctypes.data, cols.ctypes.data)
    return rows, cols[:nnz]


def optimize_bandwidth(rows, cols, perm, optimize=True, use_optimized=False):
    """OptimizeBandwidth (BandwidthOptimize.F90:182-445) with InvInitialReorder built as CreateMatrix does
    (ElementUtils.F90:1955-1958).  Returns (Perm, HalfBandWidth)."""
    rows = np.ascontiguousarray(rows, dtype=np.int32); cols = np.ascontiguousarray(cols, dtype=np.int32)
    pm = np.array(perm, dtype=np.int32)
    k = rows.size - 1
    inv = np.zeros(max(k, 1), dtype=np.int32)
    for i in np.nonzero(pm > 0)[0]:
        inv[pm[i] - 1] = i + 1
    hb = lib().orc_optimize_bandwidth(k, rows, cols, pm.size, pm, inv, int(optimize), int(use_optimized))
    return pm, hb


def initialize_matrix(rows, cols, dofs, perm_initial=None, perm=None):
    """InitializeMatrix + CRS_SortMatrix (ElementUtils.F90:1631-1732).  Returns 1-based (Rows, Cols, Diag)."""
    rows = np.ascontiguousarray(rows, dtype=np.int32); cols = np.ascontiguousarray(cols, dtype=np.int32)
    k = rows.size - 1
    nnz = int(rows[-1] - 1)
    R = np.zeros(k * dofs + 1, dtype=np.int32); Cc = np.zeros(max(nnz * dofs * dofs, 1), dtype=np.int32); D = np.zeros(max(k * dofs, 1), dtype=np.int32)
    if perm is None:
        lib().orc_initialize_matrix(k, rows, cols, dofs, None, None, R, Cc, D)
    else:
        p0 = np.ascontiguousarray(perm_initial, dtype=np.int32); p1 = np.ascontiguousarray(perm, dtype=np.int32)
        inv = np.zeros(max(k, 1), dtype=np.int32)
        for i in np.nonzero(p0 > 0)[0]:
            inv[p0[i] - 1] = i + 1
        lib().orc_initialize_matrix(k, rows, cols, dofs, p1.ctypes.data, inv.ctypes.data, R, Cc, D)
    return R, Cc[:nnz * dofs * dofs], D[:k * dofs]

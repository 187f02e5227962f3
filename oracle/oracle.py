"""TEST INFRASTRUCTURE -- ctypes front-end of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
Nothing under elmerfem_b200/ does.

The Python side only marshals arrays; every numerical loop is in elmer_oracle.cpp / fem_tools.cpp,
each citing the reference file:line it restates.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

METHODS = {"cg": 1, "bicgstab": 2, "bicgstabl": 3, "gcr": 4, "idrs": 5}
PRECONDS = {"none": 0, "diagonal": 1, "ilu0": 2, "ilu": 2}

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
    L.orc_itersolve.argtypes = [C.c_int, _ip, _ip, _ip, _dp, C.c_int, C.c_void_p, _dp, _dp, _ip, _dp,
                                C.c_int, C.c_int, C.c_void_p, _lp]
    L.orc_scale_system.restype = C.c_double
    L.orc_scale_system.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp, _dp, _dp]
    L.orc_backscale_system.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, C.c_double]
    L.orc_set_threads.argtypes = [C.c_int]
    L.orc_max_threads.restype = C.c_int
    L.fem_grid_hex8.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _ip]
    L.fem_crs_count.restype = C.c_long
    L.fem_crs_count.argtypes = [C.c_int, C.c_long, _ip, C.c_int, C.c_int, _ip]
    L.fem_crs_fill.argtypes = [C.c_int, C.c_long, _ip, C.c_int, C.c_int, _ip, _ip, _ip]
    L.fem_assemble.argtypes = [C.c_int, _dp, C.c_int, C.c_long, _ip, _dp, C.c_int, _ip, _ip, _dp, _dp, C.c_int]
    L.fem_dirichlet.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp, C.c_int, _ip, _dp, C.c_int]
    _LIB = L
    return L


def set_threads(n):
    lib().orc_set_threads(int(n))


def max_threads():
    return lib().orc_max_threads()


class CRS:
    """Elmer Matrix_t subset (Types.F90:193-283): 1-based int32 Rows/Cols/Diag, fp64 Values."""

    def __init__(self, rows, cols, diag, vals, ndeg=1):
        self.rows = np.ascontiguousarray(rows, dtype=np.int32)
        self.cols = np.ascontiguousarray(cols, dtype=np.int32)
        self.diag = np.ascontiguousarray(diag, dtype=np.int32)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.ndeg = int(ndeg)
        self.n = self.rows.size - 1
        self.nnz = self.cols.size

    def copy(self):
        return CRS(self.rows, self.cols, self.diag, self.vals.copy(), self.ndeg)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.vals, self.cols - 1, self.rows - 1), shape=(self.n, self.n))

    @staticmethod
    def from_scipy(M, ndeg=1):
        M = M.tocsr()
        M.sort_indices()
        n = M.shape[0]
        rows = (M.indptr + 1).astype(np.int32)
        cols = (M.indices + 1).astype(np.int32)
        diag = np.zeros(n, dtype=np.int32)
        for i in range(n):
            seg = cols[rows[i] - 1:rows[i + 1] - 1]
            k = np.searchsorted(seg, i + 1)
            if k >= seg.size or seg[k] != i + 1:
                raise ValueError("row %d has no diagonal entry" % (i + 1))
            diag[i] = rows[i] + k
        return CRS(rows, cols, diag, M.data.astype(np.float64), ndeg)


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
    lib().orc_crs_ilu0(A.n, A.rows, A.cols, A.diag, A.vals, ilu)
    return ilu


def lu_precond(A, ilu, v):
    u = np.empty(A.n)
    lib().orc_crs_lu_precond(A.n, A.rows, A.cols, A.diag, ilu, u, np.ascontiguousarray(v, dtype=np.float64))
    return u


def ddot(x, y):
    return lib().orc_ddot(x.size, np.ascontiguousarray(x), np.ascontiguousarray(y))


def dnrm2(x):
    return lib().orc_dnrm2(x.size, np.ascontiguousarray(x))


# ------------------------------------------------------------------ IterSolver
def fill_ipar_dpar(n, method, tol=1e-8, maxit=1000, minit=0, maxtol=1e20, residual_output=0,
                   bicgstabl_l=2, gcr_restart=None, idrs_s=4, smoothing=False, stopc=1):
    """ipar/dpar exactly as IterSolver fills them (IterSolve.F90:245-503), HUTI slots per
    fhutiter/src/huti_fdefs.h:101-155."""
    ipar = np.zeros(50, dtype=np.int32)
    dpar = np.zeros(10, dtype=np.float64)
    wrk = {"cg": 4, "bicgstab": 8}.get(method, 1)
    ipar[3 - 1] = n
    ipar[4 - 1] = wrk
    ipar[5 - 1] = residual_output
    ipar[10 - 1] = maxit
    ipar[11 - 1] = minit
    ipar[12 - 1] = stopc              # HUTI_TRESID_SCALED_BYB (IterSolve.F90:430)
    ipar[14 - 1] = 1                  # HUTI_USERSUPPLIEDX (473)
    if method == "bicgstabl":
        ipar[16 - 1] = max(2, bicgstabl_l)
    if method == "gcr":
        ipar[17 - 1] = gcr_restart if gcr_restart is not None else min(maxit, 200)
    if method == "idrs":
        ipar[18 - 1] = idrs_s
    ipar[28 - 1] = 1 if smoothing else 0
    dpar[1 - 1] = tol
    dpar[2 - 1] = maxtol
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
    if pc == 2 and ilu is None:
        ilu = ilu0(A)
    ilu_p = ilu.ctypes.data_as(C.c_void_p) if (pc == 2) else None
    Pp = None
    if method == "idrs":
        s = int(ipar[18 - 1])
        if P is None:
            P = shadow_space(n, s)
        P = np.asfortranarray(P, dtype=np.float64)
        Pp = P.ctypes.data_as(C.c_void_p)
    counts = np.zeros(4, dtype=np.int64)
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


# ------------------------------------------------------------------ synthetic problems
def grid_hex8(ex, ey, ez, lx=1.0, ly=1.0, lz=1.0):
    nn = (ex + 1) * (ey + 1) * (ez + 1)
    ne = ex * ey * ez
    xyz = np.empty(3 * nn)
    elems = np.empty(8 * ne, dtype=np.int32)
    lib().fem_grid_hex8(ex, ey, ez, lx, ly, lz, xyz, elems)
    return xyz.reshape(nn, 3), elems.reshape(ne, 8)


def crs_structure(nn, elems, ndof):
    elems = np.ascontiguousarray(elems, dtype=np.int32)
    ne, nen = elems.shape
    n = nn * ndof
    rows = np.empty(n + 1, dtype=np.int32)
    nnz = lib().fem_crs_count(nn, ne, elems.reshape(-1), nen, ndof, rows)
    if nnz < 0:
        raise OverflowError("nnz exceeds int32")
    cols = np.empty(nnz, dtype=np.int32)
    diag = np.empty(n, dtype=np.int32)
    lib().fem_crs_fill(nn, ne, elems.reshape(-1), nen, ndof, rows, cols, diag)
    return rows, cols, diag


def assemble(kind, par, xyz, elems, ndof, rows, cols, uniform=False):
    nn = xyz.shape[0]
    vals = np.empty(cols.size)
    rhs = np.empty(nn * ndof)
    p = np.zeros(8)
    p[:len(par)] = par
    elems = np.ascontiguousarray(elems, dtype=np.int32)
    lib().fem_assemble(kind, p, nn, elems.shape[0], elems.reshape(-1), np.ascontiguousarray(xyz).reshape(-1), ndof,
                       rows, cols, vals, rhs, 1 if uniform else 0)
    return vals, rhs


def dirichlet(A, b, dofs, values, symmetric=False):
    dofs = np.ascontiguousarray(dofs, dtype=np.int32)
    values = np.ascontiguousarray(np.broadcast_to(values, dofs.shape), dtype=np.float64)
    lib().fem_dirichlet(A.n, A.rows, A.cols, A.diag, A.vals, b, dofs.size, dofs, values, 1 if symmetric else 0)


def boundary_nodes(ex, ey, ez, faces="all"):
    """1-based node ids on the requested faces of the structured grid ('all' or subset of x0,x1,y0,y1,z0,z1)."""
    nx, ny, nz = ex + 1, ey + 1, ez + 1
    ids = np.arange(nx * ny * nz, dtype=np.int64).reshape(nz, ny, nx)
    sel = np.zeros_like(ids, dtype=bool)
    f = ["x0", "x1", "y0", "y1", "z0", "z1"] if faces == "all" else list(faces)
    if "x0" in f: sel[:, :, 0] = True
    if "x1" in f: sel[:, :, -1] = True
    if "y0" in f: sel[:, 0, :] = True
    if "y1" in f: sel[:, -1, :] = True
    if "z0" in f: sel[0, :, :] = True
    if "z1" in f: sel[-1, :, :] = True
    return (ids[sel] + 1).astype(np.int32)


def heat_cube(ne, faces="all", source=1.0, symmetric=False, dims=None):
    """Configs 1/2: steady heat/Poisson on the unit cube, ne^3 hex8, u=0 on `faces`, f=source."""
    ex, ey, ez = dims if dims is not None else (ne, ne, ne)
    xyz, elems = grid_hex8(ex, ey, ez)
    rows, cols, diag = crs_structure(xyz.shape[0], elems, 1)
    vals, rhs = assemble(0, [source], xyz, elems, 1, rows, cols, uniform=True)
    A = CRS(rows, cols, diag, vals, 1)
    dirichlet(A, rhs, boundary_nodes(ex, ey, ez, faces), 0.0, symmetric)
    return A, rhs


def elasticity_beam(ex, ey, ez, lx=8.0, ly=1.0, lz=1.0, E=1e9, nu=0.3, load=(0.0, 0.0, -1e4)):
    """Configs 3/5: isotropic linear elasticity, 3 interleaved dofs/node, x=0 end clamped, body load."""
    xyz, elems = grid_hex8(ex, ey, ez, lx, ly, lz)
    rows, cols, diag = crs_structure(xyz.shape[0], elems, 3)
    vals, rhs = assemble(1, [E, nu, load[0], load[1], load[2]], xyz, elems, 3, rows, cols, uniform=True)
    A = CRS(rows, cols, diag, vals, 3)
    nodes = boundary_nodes(ex, ey, ez, ["x0"]).astype(np.int64)
    dofs = np.concatenate([3 * (nodes - 1) + c + 1 for c in range(3)]).astype(np.int32)
    dofs.sort()
    dirichlet(A, rhs, dofs, 0.0, False)
    return A, rhs


def cavity_flow(ne, visc=0.01, tau=None):
    """Config 4: nonsymmetric 4-dof/node (u,v,w,p) Picard-linearised stabilised system on the unit cube;
    velocity fixed on all walls (lid z=1 moving in x), pressure pinned at node 1."""
    xyz, elems = grid_hex8(ne, ne, ne)
    rows, cols, diag = crs_structure(xyz.shape[0], elems, 4)
    if tau is None:
        tau = 0.5 / ne
    vals, rhs = assemble(2, [visc, tau, 0.0], xyz, elems, 4, rows, cols, uniform=False)
    A = CRS(rows, cols, diag, vals, 4)
    wall = boundary_nodes(ne, ne, ne, "all").astype(np.int64)
    lid = set(boundary_nodes(ne, ne, ne, ["z1"]).tolist())
    dofs, dv = [], []
    for nd in wall:
        for c in range(3):
            dofs.append(4 * (nd - 1) + c + 1)
            dv.append(1.0 if (c == 0 and nd in lid) else 0.0)
    dofs.append(4)
    dv.append(0.0)
    dofs = np.array(dofs, dtype=np.int32)
    dv = np.array(dv)
    o = np.argsort(dofs)
    dirichlet(A, rhs, dofs[o], dv[o], False)
    return A, rhs

"""Synthetic workloads of BASELINE.json (host tool, not on the solve path).

Structured hex8 meshes numbered as ElmerGrid numbers them, the CRS structure CreateMatrix builds for
nodal dofs, hex8 element assembly, Dirichlet rows and Elmer's default diagonal scaling -- i.e. the
arrays Elmer would pass to IterSolver for the configs: heat cube (C1/C2), elasticity beam (C3/C5),
lid-driven cavity block system (C4).  All numerical loops are in fem_tools.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-C", _HERE, "libelmer_synth.so"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "libelmer_synth.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.fem_grid_hex8.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _ip]
    L.fem_crs_count.restype = C.c_long
    L.fem_crs_count.argtypes = [C.c_int, C.c_long, _ip, C.c_int, C.c_int, _ip]
    L.fem_crs_fill.argtypes = [C.c_int, C.c_long, _ip, C.c_int, C.c_int, _ip, _ip, _ip]
    L.fem_assemble.argtypes = [C.c_int, _dp, C.c_int, C.c_long, _ip, _dp, C.c_int, _ip, _ip, _dp, _dp, C.c_int]
    L.fem_dirichlet.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp, C.c_int, _ip, _dp, C.c_int]
    L.fem_scale_system.restype = C.c_double
    L.fem_scale_system.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp, _dp]
    _LIB = L
    return L


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


def scale_system(A, b):
    """In place `Linear System Scaling` (ScaleLinearSystemDiagonal, SolverUtils.F90:12976-13213);
    returns (D, bnorm) with x_unscaled = D * x_scaled."""
    D = np.zeros(A.n)
    bn = lib().fem_scale_system(A.n, A.rows, A.cols, A.diag, A.vals, b, D)
    return D, bn


def workload(name, ne=None):
    """The BASELINE.json configs as scaled systems (A, b, ndeg).  ne overrides the element count per
    edge (tests use small ne; bench.py the full sizes)."""
    if name == "heat":
        A, b = heat_cube(ne or 200, faces="all", source=1.0)
    elif name == "elasticity":
        e = ne or 136
        A, b = elasticity_beam(e, e, e, lx=1.0)
    elif name == "cavity":
        A, b = cavity_flow(ne or 135)
    else:
        raise ValueError(name)
    scale_system(A, b)
    return A, b


def slab_layers(nz_nodes, nranks):
    """Node layers [L[r], L[r+1]) owned by rank r: contiguous z-slabs, i.e. contiguous global rows in
    ElmerGrid's x-fastest numbering (the continuous numbering of SParIterSolver.F90:1453-1488 is then
    the natural one)."""
    return [(r * nz_nodes) // nranks for r in range(nranks + 1)]


def heat_slab(ex, ey, ez_total, rank, nranks, allreduce_sum=None, source=1.0):
    """Rank `rank`'s share of the heat problem on an ex x ey x ez_total hex8 box (cubic elements of
    side 1/ex), Dirichlet u=0 on the whole boundary, Elmer's diagonal scaling applied: complete owned
    rows with GLOBAL 1-based column ids, as b200_set_partition expects.  Only a slab two element
    layers thicker than the owned one is ever assembled.  allreduce_sum(float) sums over ranks (needed
    for the global ||D b|| of the scaling); None = single process emulation of all ranks is not done here."""
    return _slab("heat", ex, ey, ez_total, rank, nranks, allreduce_sum, source=source)


def elasticity_slab(ex, ey, ez_total, rank, nranks, allreduce_sum=None, E=1e9, nu=0.3, load=(0.0, -1e4, 0.0)):
    """Configs 3/5: rank `rank`'s share of a linear-elasticity beam (3 interleaved dofs per node, long axis z,
    cubic hex8 elements of side 1/ex, z=0 end clamped, body load), partitioned into z-slabs of node layers
    (`ElmerGrid -partition 1 1 N`); same conventions as heat_slab."""
    return _slab("elasticity", ex, ey, ez_total, rank, nranks, allreduce_sum, E=E, nu=nu, load=load)


def _slab(kind, ex, ey, ez_total, rank, nranks, allreduce_sum, source=1.0, E=1e9, nu=0.3, load=(0.0, -1e4, 0.0)):
    nx, ny, NZ = ex + 1, ey + 1, ez_total + 1
    L = slab_layers(NZ, nranks)
    zlo, zhi = max(0, L[rank] - 2), min(NZ, L[rank + 1] + 2)        # local node layers [zlo, zhi)
    ezl = zhi - zlo - 1
    h = 1.0 / ex
    xyz, elems = grid_hex8(ex, ey, ezl, ex * h, ey * h, ezl * h)
    nd = 1 if kind == "heat" else 3
    rows, cols, diag = crs_structure(xyz.shape[0], elems, nd)
    if kind == "heat":
        vals, rhs = assemble(0, [source], xyz, elems, 1, rows, cols, uniform=True)
        A = CRS(rows, cols, diag, vals, 1)
        faces = ["x0", "x1", "y0", "y1"] + (["z0"] if zlo == 0 else []) + (["z1"] if zhi == NZ else [])
        dirichlet(A, rhs, boundary_nodes(ex, ey, ezl, faces), 0.0, False)
    else:
        vals, rhs = assemble(1, [E, nu, load[0], load[1], load[2]], xyz, elems, 3, rows, cols, uniform=True)
        A = CRS(rows, cols, diag, vals, 3)
        if zlo == 0:
            nodes = boundary_nodes(ex, ey, ezl, ["z0"]).astype(np.int64)
            dofs = np.concatenate([3 * (nodes - 1) + c + 1 for c in range(3)]).astype(np.int32)
            dofs.sort()
            dirichlet(A, rhs, dofs, 0.0, False)
    # scaling with D from the complete rows (the two outermost artificial layers are never referenced)
    d = np.abs(A.vals[A.diag - 1])
    D = 1.0 / np.sqrt(d)
    plane = nx * ny * nd                                            # dofs per node layer
    o0, o1 = plane * (L[rank] - zlo), plane * (L[rank + 1] - zlo)
    p0, p1 = A.rows[o0] - 1, A.rows[o1] - 1
    lrows = (A.rows[o0:o1 + 1] - A.rows[o0] + 1).astype(np.int32)
    lcols = A.cols[p0:p1]
    rowid = np.repeat(np.arange(o0, o1, dtype=np.int64), np.diff(A.rows[o0:o1 + 1]))
    lvals = A.vals[p0:p1] * (D[rowid] * D[lcols - 1])
    b = rhs[o0:o1] * D[o0:o1]
    s = float(np.dot(b, b))
    if allreduce_sum is not None:
        s = allreduce_sum(s)
    bnorm = np.sqrt(s)
    b = b / bnorm
    gcols = (lcols.astype(np.int64) + plane * zlo).astype(np.int32)
    goffset = np.array([plane * l for l in L], dtype=np.int32)
    return dict(rows=np.ascontiguousarray(lrows), cols=np.ascontiguousarray(gcols), vals=np.ascontiguousarray(lvals),
                b=np.ascontiguousarray(b), goffset=goffset, gn=int(plane * NZ), D=D[o0:o1] * bnorm, bnorm=bnorm, ndeg=nd)

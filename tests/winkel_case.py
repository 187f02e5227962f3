"""The reference's partitioned benchmark cases fem/tests/WinkelBmPoisson{CgIlu0,IdrsIlu0}: -div(grad u) = 1 (StatCurrentSolver, conductivity 1,
source 1) on the `winkel` hex8 mesh, u = 0 on boundary 1, `Reference Norm = 1.03281284` at every partition count (case.sif).  The mesh is
generated here from the reference's winkel.grd (copied to tests/golden/winkel/) with the reference's own ElmerGrid (oracle/_ref/ElmerGrid, built
by `make -C oracle ref`), partitioned with `-partdual -metisrec N` exactly as the test's runtest.cmake does."""
import os
import shutil
import subprocess
import tempfile

import numpy as np

from elmerfem_b200 import meshio, synth

HERE = os.path.dirname(os.path.abspath(__file__))
ELMERGRID = os.path.join(HERE, "..", "oracle", "_ref", "ElmerGrid")
GRD = os.path.join(HERE, "golden", "winkel", "winkel.grd")
REFERENCE_NORM = 1.03281284
_cache = {}


def available():
    return os.path.exists(ELMERGRID) and os.access(ELMERGRID, os.X_OK)


def mesh_dir(nparts=0):
    """Directory holding mesh.* (and partitioning.N for nparts > 0), generated once per process."""
    if "dir" not in _cache:
        d = tempfile.mkdtemp(prefix="winkel_")
        shutil.copy(GRD, os.path.join(d, "winkel.grd"))
        subprocess.check_call([ELMERGRID, "1", "2", "winkel"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _cache["dir"] = d
    d = _cache["dir"]
    if nparts > 0 and not os.path.isdir(os.path.join(d, "winkel", "partitioning.%d" % nparts)):
        subprocess.check_call([ELMERGRID, "1", "2", "winkel", "-partdual", "-metisrec", str(nparts), "-nooverwrite"], cwd=d,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.join(d, "winkel")


def system():
    """(A, b): assembled, Dirichlet rows set, NOT scaled; natural (ElmerGrid) node numbering."""
    if "sys" not in _cache:
        m = meshio.read_mesh(mesh_dir())
        nid = np.zeros(int(m.node_ids.max()) + 1, dtype=np.int64)
        nid[m.node_ids] = np.arange(m.node_ids.size)
        xyz = np.ascontiguousarray(m.xyz)
        elems = np.ascontiguousarray(np.array([nid[c] for c in m.elems], dtype=np.int32) + 1)
        rows, cols, diag = synth.crs_structure(xyz.shape[0], elems, 1)
        vals, rhs = synth.assemble(0, [1.0], xyz, elems, 1, rows, cols, uniform=False)
        A = synth.CRS(rows, cols, diag, vals, 1)
        fixed = np.array(sorted({int(nid[g]) + 1 for b in m.bnd if b[1] == 1 for g in b[5]}), dtype=np.int32)
        synth.dirichlet(A, rhs, fixed, 0.0, False)
        _cache["sys"] = (A, rhs)
    A, rhs = _cache["sys"]
    return A.copy(), rhs.copy()


def norm(x):
    return float(np.sqrt(np.sum(x * x) / x.size))


# ---------------------------------------------------------------------------------------------------------------------------------
# fem/tests/WinkelBmNavier* (case.sif identical across the family): StressSolver, 3 dofs per node, E = 1e9, nu = 0.3, wall (boundary 3)
# clamped, traction Force 2 = 1e6 on boundary 7, `Reference Norm = 2.25252433E-02`.  Coarser winkel.grd (Reference Density 0.25).
NAVIER_GRD = os.path.join(HERE, "golden", "winkel", "winkel_navier.grd")
NAVIER_REFERENCE_NORM = 2.25252433E-02


def navier_mesh_dir(nparts=0, method="-metiskway"):
    if "ndir" not in _cache:
        d = tempfile.mkdtemp(prefix="winkel_navier_")
        shutil.copy(NAVIER_GRD, os.path.join(d, "winkel.grd"))
        subprocess.check_call([ELMERGRID, "1", "2", "winkel"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        _cache["ndir"] = d
    d = _cache["ndir"]
    if nparts > 0 and not os.path.isdir(os.path.join(d, "winkel", "partitioning.%d" % nparts)):
        subprocess.check_call([ELMERGRID, "1", "2", "winkel", "-partdual", method, str(nparts), "-nooverwrite"], cwd=d,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return os.path.join(d, "winkel")


def navier_system():
    """(A, b) of the elasticity case: ndeg = 3, natural numbering, not scaled."""
    if "nsys" not in _cache:
        m = meshio.read_mesh(navier_mesh_dir())
        nid = np.zeros(int(m.node_ids.max()) + 1, dtype=np.int64)
        nid[m.node_ids] = np.arange(m.node_ids.size)
        xyz = np.ascontiguousarray(m.xyz)
        elems = np.ascontiguousarray(np.array([nid[c] for c in m.elems], dtype=np.int32) + 1)
        rows, cols, diag = synth.crs_structure(xyz.shape[0], elems, 3)
        vals, rhs = synth.assemble(1, [1.0e9, 0.3, 0.0, 0.0, 0.0], xyz, elems, 3, rows, cols, uniform=False)
        A = synth.CRS(rows, cols, diag, vals, 3)
        g = 1.0 / np.sqrt(3.0)                                       # surface traction: bilinear quads, 2 x 2 Gauss
        for bnd in m.bnd:
            if bnd[1] != 7:
                continue
            q = [int(nid[v]) for v in bnd[5]]
            Pq = xyz[q]
            for xi in (-g, g):
                for eta in (-g, g):
                    N = 0.25 * np.array([(1 - xi) * (1 - eta), (1 + xi) * (1 - eta), (1 + xi) * (1 + eta), (1 - xi) * (1 + eta)])
                    dxi = 0.25 * np.array([-(1 - eta), (1 - eta), (1 + eta), -(1 + eta)])
                    deta = 0.25 * np.array([-(1 - xi), -(1 + xi), (1 + xi), (1 - xi)])
                    J = np.linalg.norm(np.cross(dxi @ Pq, deta @ Pq))
                    for a in range(4):
                        rhs[3 * q[a] + 1] += 1.0e6 * N[a] * J
        nodes = np.array(sorted({int(nid[v]) for bnd in m.bnd if bnd[1] == 3 for v in bnd[5]}), dtype=np.int64)
        dofs = np.sort(np.concatenate([3 * nodes + c + 1 for c in range(3)])).astype(np.int32)
        synth.dirichlet(A, rhs, dofs, 0.0, False)
        _cache["nsys"] = (A, rhs)
    A, rhs = _cache["nsys"]
    return A.copy(), rhs.copy()

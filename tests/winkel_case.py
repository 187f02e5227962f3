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

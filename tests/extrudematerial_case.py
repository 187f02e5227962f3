"""The reference's fem/tests/ElmerGridExtrudeMaterial case: HeatSolver without source on the hex8 mesh ElmerGrid extrudes
from cubes.grd (two bodies, Heat Conductivity 1 and 2), Temperature = 0 on boundaries 101, 102 and 1 on 501..504,
BiCGStab + ILU0 at 1e-8.  case.sif: `Solver 1 :: Reference Norm = Real 0.67120112`.
The mesh is made at test time by the reference's own ElmerGrid (oracle/_ref/ElmerGrid), as runtest.cmake does."""
import os
import shutil
import subprocess
import tempfile

import numpy as np

import elmerfem_b200 as b200
from elmerfem_b200 import meshio, synth

HERE = os.path.dirname(os.path.abspath(__file__))
ELMERGRID = os.path.join(HERE, "..", "oracle", "_ref", "ElmerGrid")
GRD = os.path.join(HERE, "golden", "extrudematerial", "cubes.grd")
REFERENCE_NORM = 0.67120112
CONDUCTIVITY = {1: 1.0, 2: 2.0}
COLD, HOT = (101, 102), (501, 502, 503, 504)
_cache = {}


def available():
    return os.path.exists(ELMERGRID) and os.access(ELMERGRID, os.X_OK)


def system():
    """(A, b, perm): assembled in the numbering CreateMatrix gives the mesh, Dirichlet rows set, not scaled."""
    if "sys" not in _cache:
        d = tempfile.mkdtemp(prefix="extrude_")
        shutil.copy(GRD, os.path.join(d, "cubes.grd"))
        subprocess.check_call([ELMERGRID, "1", "2", "cubes.grd"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        m = meshio.read_mesh(os.path.join(d, "cubes"))
        shutil.rmtree(d, ignore_errors=True)
        nid = np.zeros(int(m.node_ids.max()) + 1, dtype=np.int64)
        nid[m.node_ids] = np.arange(1, m.node_ids.size + 1)
        nn = m.node_ids.size
        bulk = np.array([nid[c] for c in m.elems], dtype=np.int32)
        bnd = [nid[t[5]].astype(np.int32) for t in m.bnd]
        flat = np.concatenate([bulk.reshape(-1)] + bnd).astype(np.int32)
        ptr = np.zeros(len(bulk) + len(bnd) + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([8] * len(bulk) + [len(e) for e in bnd])
        S = b200.create_matrix_structure(ptr, flat, nn, dofs=1)
        perm = S["perm"]
        xyz = np.empty_like(m.xyz); xyz[perm - 1] = m.xyz            # coordinates in matrix numbering
        vals = np.zeros(S["cols"].size)
        for body, k in CONDUCTIVITY.items():
            el = np.ascontiguousarray(perm[bulk[np.asarray(m.elem_body) == body] - 1], dtype=np.int32)
            v, _ = synth.assemble(0, [0.0], np.ascontiguousarray(xyz), el, 1, S["rows"], S["cols"], uniform=False)
            vals += k * v
        A = synth.CRS(S["rows"], S["cols"], S["diag"], vals, 1)
        rhs = np.zeros(nn)
        for tags, value in ((COLD, 0.0), (HOT, 1.0)):
            nodes = sorted({int(perm[g - 1]) for t, e in zip(m.bnd, bnd) if t[1] in tags for g in e})
            synth.dirichlet(A, rhs, np.array(nodes, dtype=np.int32), value, False)
        _cache["sys"] = (A, rhs, perm)
    A, rhs, perm = _cache["sys"]
    return A.copy(), rhs.copy(), perm


def compute_norm(x):
    return float(np.sqrt(np.sum(x * x) / x.size))

"""The reference's fem/tests/CoordinateScaling case as a system the tests can solve: HeatSolver (conductivity 1, heat
source 1, density 1) on the 20 x 20 bilinear-quad mesh ElmerGrid makes from square.grd (tests/golden/coordinatescaling =
`ElmerGrid 1 2 square` run with the reference's own ElmerGrid, as runtest.cmake does), `Coordinate Scaling = 0.001`
(the 1000 x 1000 square becomes the unit square), Temperature = 0 on boundary 1, BiCGStab + ILU1 at 1e-8.
case.sif: `Solver 1 :: Reference Norm = Real 3.93779036434094704E-002`.

The matrix structure and the numbering come from the library's structure producer (CreateMatrix's nodal path with the
default `Optimize Bandwidth = True`), so the ILU(1) factor is built in the order Elmer builds it in."""
import os

import numpy as np
import scipy.sparse as sp

import elmerfem_b200 as b200
from elmerfem_b200 import meshio, synth

MESH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coordinatescaling")
REFERENCE_NORM = 3.93779036434094704E-002
_G = 1.0 / np.sqrt(3.0)


def _quad_element(p):
    """Bilinear quad: K = int grad N . grad N, f = int N (2 x 2 Gauss, exact on a parallelogram)."""
    K = np.zeros((4, 4)); f = np.zeros(4)
    for xi in (-_G, _G):
        for eta in (-_G, _G):
            N = 0.25 * np.array([(1 - xi) * (1 - eta), (1 + xi) * (1 - eta), (1 + xi) * (1 + eta), (1 - xi) * (1 + eta)])
            dN = 0.25 * np.array([[-(1 - eta), (1 - eta), (1 + eta), -(1 + eta)],
                                  [-(1 - xi), -(1 + xi), (1 + xi), (1 - xi)]])
            J = dN @ p
            det = J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]
            g = np.linalg.solve(J, dN)
            K += (g.T @ g) * det
            f += N * det
    return K, f


def system(optimize_bandwidth=True):
    """(A, b, info): CRS in the numbering CreateMatrix gives this mesh, Dirichlet rows set (identity, structure kept)."""
    m = meshio.read_mesh(MESH)
    nid = np.zeros(int(m.node_ids.max()) + 1, dtype=np.int64)
    nid[m.node_ids] = np.arange(1, m.node_ids.size + 1)
    nn = m.node_ids.size
    xy = m.xyz[:, :2] * 0.001                                         # Coordinate Scaling
    bulk = [nid[c] for c in m.elems]
    bnd = [nid[t[5]] for t in m.bnd]
    allel = bulk + bnd                                                # Mesh % Elements: bulk, then boundary
    ptr = np.zeros(len(allel) + 1, dtype=np.int32); ptr[1:] = np.cumsum([len(e) for e in allel])
    S = b200.create_matrix_structure(ptr, np.concatenate(allel).astype(np.int32), nn, dofs=1, optimize_bw=optimize_bandwidth)
    perm = S["perm"]
    I, J, V = [], [], []
    rhs = np.zeros(nn)
    for e in bulk:
        K, f = _quad_element(xy[e - 1])
        r = perm[e - 1] - 1
        for a in range(4):
            rhs[r[a]] += f[a]
            for c in range(4):
                I.append(r[a]); J.append(r[c]); V.append(K[a, c])
    M = sp.csr_matrix((V, (I, J)), shape=(nn, nn)); M.sum_duplicates(); M.sort_indices()
    assert np.array_equal(M.indptr + 1, S["rows"]) and np.array_equal(M.indices + 1, S["cols"])   # producer == assembled pattern
    fixed = sorted({int(perm[g - 1]) - 1 for t, e in zip(m.bnd, bnd) if t[1] == 1 for g in e})
    for i in fixed:
        lo, hi = M.indptr[i], M.indptr[i + 1]
        M.data[lo:hi] = np.where(M.indices[lo:hi] == i, 1.0, 0.0)
        rhs[i] = 0.0
    A = synth.CRS(S["rows"], S["cols"], S["diag"], M.data, 1)
    return A, rhs, dict(perm=perm, half_bandwidth=S["half_bandwidth"], nn=nn)


def compute_norm(x):
    """ComputeNorm, default: sqrt(sum x^2 / n) (SolverUtils.F90:10290-...)."""
    return float(np.sqrt(np.sum(x * x) / x.size))

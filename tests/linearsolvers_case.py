"""The reference's own linear-solver regression case, fem/tests/linearsolvers/TempDist.sif, as a system the tests can solve:
HeatSolver (Laplace, conductivity 1, no source) on the 2-D triangle mesh of that test (tests/golden/linearsolvers = a copy of
fem/tests/linearsolvers/Mesh), Temperature = k on boundaries 1..6, natural conditions elsewhere.  The exact answer is the
constant k, so the reference checks `Reference Norm = k` for every solver (TempDist.sif:52-175)."""
import os

import numpy as np
import scipy.sparse as sp

from elmerfem_b200 import meshio

MESH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "linearsolvers")


def tempdist_system(k):
    """(S, b, x0): P1-triangle stiffness matrix with Dirichlet rows T = k on boundaries 1..6 (row -> identity, structure kept);
    x0 = the initial guess Elmer starts from: zero with the Dirichlet values already set in it."""
    m = meshio.read_mesh(MESH)
    nid = {int(g): i for i, g in enumerate(m.node_ids)}
    n = len(nid)
    xy = m.xyz[:, :2]
    I, J, V = [], [], []
    for conn in m.elems:
        t = [nid[int(g)] for g in conn]
        p = xy[t]
        B = np.array([[p[1, 1] - p[2, 1], p[2, 1] - p[0, 1], p[0, 1] - p[1, 1]],
                      [p[2, 0] - p[1, 0], p[0, 0] - p[2, 0], p[1, 0] - p[0, 0]]])
        area = 0.5 * abs((p[1, 0] - p[0, 0]) * (p[2, 1] - p[0, 1]) - (p[2, 0] - p[0, 0]) * (p[1, 1] - p[0, 1]))
        Ke = (B.T @ B) / (4.0 * area)
        for a in range(3):
            for c in range(3):
                I.append(t[a]); J.append(t[c]); V.append(Ke[a, c])
    S = sp.csr_matrix((V, (I, J)), shape=(n, n))
    S.sum_duplicates()
    S.sort_indices()
    b = np.zeros(n)
    fixed = sorted({nid[int(g)] for (_, bc, _, _, _, nodes) in m.bnd if 1 <= bc <= 6 for g in nodes})
    for i in fixed:                                   # Dirichlet row: zeros stay in the structure, as in Elmer
        lo, hi = S.indptr[i], S.indptr[i + 1]
        S.data[lo:hi] = np.where(S.indices[lo:hi] == i, 1.0, 0.0)
        b[i] = float(k)
    x0 = np.zeros(n)
    x0[fixed] = float(k)
    return S, b, x0


def compute_norm(x):
    """ComputeNorm, default 2-norm scaled by the number of dofs (SolverUtils.F90:10290-...): sqrt(sum x^2 / n)."""
    return float(np.sqrt(np.sum(x * x) / x.size))

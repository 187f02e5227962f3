"""Property tests (hypothesis) of the integer work and of the oracle's factorisations on random structurally symmetric matrices:
the cases the structured meshes never produce (irregular rows, arbitrary contiguous partitions, ranks without neighbours)."""
import numpy as np
import scipy.sparse as sp
from hypothesis import given, settings, strategies as st

from elmerfem_b200 import synth


def random_sym_matrix(n, density, seed):
    rs = np.random.RandomState(seed)
    m = int(max(1, density * n * n / 2))
    i = rs.randint(0, n, m); j = rs.randint(0, n, m)
    v = rs.standard_normal(m)
    S = sp.coo_matrix((v, (i, j)), shape=(n, n))
    S = (S + S.T).tocsr()
    S = S + sp.diags(np.abs(S).sum(axis=1).A.ravel() + 1.0)          # diagonally dominant, full diagonal
    S = S.tocsr(); S.sort_indices()
    return S


@settings(max_examples=25, deadline=None)
@given(n=st.integers(5, 120), density=st.floats(0.01, 0.3), seed=st.integers(0, 10 ** 6), nparts=st.integers(1, 6))
def test_halo_plan_random_partitions(b200, n, density, seed, nparts):
    """Send lists, ghost slots, owned/ghost split and the peer-memory layout for arbitrary contiguous row partitions of a random
    structurally symmetric matrix: bit-exact against the restatement of the reference's elmer_distribute_matrix."""
    from test_halo_plan import check_against_oracle
    S = random_sym_matrix(n, density, seed)
    rs = np.random.RandomState(seed + 1)
    cuts = np.sort(rs.choice(np.arange(1, n), size=min(nparts - 1, n - 1), replace=False)) if nparts > 1 else np.zeros(0, dtype=int)
    goffset = np.concatenate([[0], cuts, [n]]).astype(np.int32)
    check_against_oracle(b200, S, goffset)


def dense_ilu_on_pattern(A, P):
    n = A.shape[0]
    LU = A.copy()
    for i in range(n):
        for k in range(i):
            if P[i, k] and LU[i, k] != 0.0:
                LU[i, k] /= LU[k, k]
                for j in range(k + 1, n):
                    if P[i, j]:
                        LU[i, j] -= LU[i, k] * LU[k, j]
    return LU


@settings(max_examples=20, deadline=None)
@given(n=st.integers(4, 60), density=st.floats(0.02, 0.25), seed=st.integers(0, 10 ** 6), order=st.integers(0, 2))
def test_ilun_random_matrices(oracle, n, density, seed, order):
    """ILU(n) of the oracle (pattern by rounds of InitializeILU1, values by the row-wise elimination) against a set-based pattern
    construction and a dense IKJ elimination on random matrices; L U (M^-1 v) = v."""
    S = random_sym_matrix(n, density, seed)
    A = synth.CRS.from_scipy(S)
    F = oracle.ilun(A, order) if order else synth.CRS(A.rows, A.cols, A.diag, oracle.ilu0(A), 1)
    pat = [set((A.cols[A.rows[i] - 1:A.rows[i + 1] - 1] - 1).tolist()) for i in range(n)]
    for _ in range(order):
        pat = [set(pat[i]) | {j for k in pat[i] if k < i for j in pat[k] if j > k} for i in range(n)]
    for i in range(n):
        assert sorted(pat[i]) == (F.cols[F.rows[i] - 1:F.rows[i + 1] - 1] - 1).tolist()
    P = np.zeros((n, n), dtype=bool)
    for i in range(n):
        P[i, sorted(pat[i])] = True
    LU = dense_ilu_on_pattern(S.toarray(), P)
    r0 = np.repeat(np.arange(n), np.diff(F.rows))
    vals = F.vals.copy(); vals[F.diag - 1] = 1.0 / vals[F.diag - 1]
    assert np.abs(vals - LU[r0, F.cols - 1]).max() <= 1e-10 * max(1.0, np.abs(LU).max())
    v = np.random.RandomState(seed + 2).standard_normal(n)
    u = oracle.lu_precond(A, F if order else F.vals, v)
    back = (np.tril(LU, -1) * P + np.eye(n)) @ ((np.triu(LU) * P) @ u)
    assert np.abs(back - v).max() <= 1e-9 * max(1.0, np.abs(v).max())


@settings(max_examples=40, deadline=None)
@given(nn=st.integers(2, 150), ne=st.integers(1, 200), seed=st.integers(0, 10 ** 6), frac=st.floats(0.3, 1.0), dofs=st.integers(1, 3),
       use_optimized=st.booleans())
def test_create_matrix_random_meshes(oracle, b200, nn, ne, seed, frac, dofs, use_optimized):
    """CreateMatrix's nodal path (node graph -> OptimizeBandwidth -> InitializeMatrix) of the library against the list-matrix restatement,
    on random "meshes": elements of 1..8 nodes, equations that cover only part of the nodes, isolated nodes, several components."""
    rs = np.random.RandomState(seed)
    elems = [rs.choice(nn, size=rs.randint(1, min(8, nn) + 1), replace=False) + 1 for _ in range(ne)]
    ptr = np.zeros(ne + 1, dtype=np.int32); ptr[1:] = np.cumsum([len(e) for e in elems])
    nodes = np.concatenate(elems).astype(np.int32)
    appears = np.zeros(nn, dtype=bool); appears[nodes - 1] = True
    active = (rs.rand(nn) < frac) & appears                    # nodes outside every element carry no dof of the equation
    active[nodes[0] - 1] = True
    k = int(active.sum())
    perm0 = np.zeros(nn, dtype=np.int32)
    perm0[rs.permutation(np.nonzero(active)[0])] = np.arange(1, k + 1)
    lr_o, lc_o = oracle.make_list_matrix(ptr, nodes, perm0, k)
    lr_p, lc_p = b200.node_graph(ptr, nodes, nn, perm0, k)
    assert np.array_equal(lr_p, lr_o) and np.array_equal(lc_p, lc_o)
    assert np.all(np.diff(lr_o) > 0)
    p_o, hb_o = oracle.optimize_bandwidth(lr_o, lc_o, perm0, True, use_optimized)
    p_p, hb_p = b200.optimize_bandwidth(lr_p, lc_p, perm0, True, use_optimized)
    assert hb_p == hb_o and np.array_equal(p_p, p_o)
    assert np.array_equal(np.sort(p_p[active]), np.arange(1, k + 1)) and np.all(p_p[~active] == 0)
    got = b200.initialize_structure(lr_p, lc_p, dofs, perm0, p_p)
    ref = oracle.initialize_matrix(lr_o, lc_o, dofs, perm0, p_o)
    for g, r in zip(got, ref):
        assert np.array_equal(g, r)
    R, Cc, D = got
    assert np.array_equal(Cc[D - 1], np.arange(1, k * dofs + 1))          # Diag points at the diagonal
    assert all(np.all(np.diff(Cc[R[i] - 1:R[i + 1] - 1]) > 0) for i in range(k * dofs))

"""Incomplete Cholesky ('Linear System Symmetric ILU', A % Cholesky): the Cholesky branches of CRS_IncompleteLU (fem/src/CRSMatrix.F90:
3539-3602) and CRS_LUSolve (4618-4638).  The reference holds no test case with this keyword (parity unpinned by a golden vector): the
oracle's restatement is checked against an independent dense IC(0) / IC(1) and, on the GPU, the library must reproduce the restatement bit
for bit (factor, both sweeps) and take its iteration counts."""
import numpy as np
import pytest
import scipy.sparse as sp


@pytest.fixture()
def cholesky(oracle):
    oracle.set_cholesky(True)
    yield
    oracle.set_cholesky(False)


def _spd(oracle, dims=(6, 5, 4)):
    A, b = oracle.heat_cube(0, faces=["x0"], dims=dims, symmetric=True)
    return A, b


def _dense_ic(M, pat):
    n = M.shape[0]
    L = np.zeros((n, n))
    for i in range(n):
        for j in range(i):
            if pat[i, j]:
                L[i, j] = (M[i, j] - L[i, :j] @ L[j, :j]) / L[j, j]
        L[i, i] = np.sqrt(M[i, i] - L[i, :i] @ L[i, :i])
    return L


def _lower(F):
    M = sp.csr_matrix((F.vals, F.cols - 1, F.rows - 1), shape=(F.n, F.n)).toarray()
    return np.tril(M, -1) + np.diag(1.0 / np.diag(M))


def test_oracle_ic0_and_ic1_against_dense_factorisation(oracle, cholesky):
    A, b = _spd(oracle)
    M = sp.csr_matrix((A.vals, A.cols - 1, A.rows - 1), shape=(A.n, A.n)).toarray()
    assert np.abs(M - M.T).max() == 0.0
    F0 = A.copy(); F0.vals = oracle.ilu0(A)
    assert np.abs(_lower(F0) - _dense_ic(M, M != 0)).max() < 1e-14
    assert np.all(np.triu(sp.csr_matrix((F0.vals, F0.cols - 1, F0.rows - 1), shape=(A.n, A.n)).toarray(), 1) == 0.0)     # upper part unwritten (here 0)
    F1 = oracle.ilun(A, 1)
    pat1 = sp.csr_matrix((np.ones(F1.cols.size), F1.cols - 1, F1.rows - 1), shape=(A.n, A.n)).toarray() != 0
    assert np.abs(_lower(F1) - _dense_ic(M, pat1)).max() < 1e-14
    v = np.random.RandomState(0).standard_normal(A.n)
    for F, pat in ((F0, M != 0), (F1, pat1)):
        L = _dense_ic(M, pat)
        ref = np.linalg.solve(L.T, np.linalg.solve(L, v))
        got = oracle.lu_precond(A, F.vals, v) if F is F0 else oracle.lu_precond(A, F, v)
        assert np.abs(got - ref).max() < 1e-13 * np.abs(ref).max()


def test_oracle_cg_with_incomplete_cholesky_converges_like_ilu0_on_spd(oracle, cholesky):
    """On an SPD matrix IC(0) and ILU(0) are the same preconditioner in exact arithmetic (U = D L^T): same CG iteration count."""
    A, b = _spd(oracle, (10, 10, 10))
    r1 = oracle.itersolve(A, b, method="cg", precond="ilu0", tol=1e-10, maxit=300)
    oracle.set_cholesky(False)
    r0 = oracle.itersolve(A, b, method="cg", precond="ilu0", tol=1e-10, maxit=300)
    assert r1["info"] == r0["info"] == 1 and abs(r1["iters"] - r0["iters"]) <= 1
    assert np.abs(r1["x"] - r0["x"]).max() < 1e-8 * np.abs(r0["x"]).max()


def test_negative_pivot_takes_the_reference_guard(oracle, cholesky):
    """S(i) <= AEPS stores 1 (CRSMatrix.F90:3578-3585)."""
    A, b = _spd(oracle, (3, 3, 2))
    A = A.copy(); A.vals = A.vals.copy(); A.vals[A.diag[5] - 1] = -1.0
    ic = oracle.ilu0(A)
    assert ic[A.diag[5] - 1] == 1.0 and np.all(np.isfinite(ic))


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dims,order", [((6, 5, 4), 0), ((9, 9, 9), 0), ((20, 7, 3), 0), ((6, 5, 4), 1), ((8, 8, 8), 2)])
def test_gpu_factor_and_solve_bit_exact(oracle, b200, cholesky, dims, order):
    A, b = _spd(oracle, dims)
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, 1)
        M.set_values(A.vals)
        M.set_symmetric_ilu(True)
        if order:
            M.set_ilu_order(order)
        M.factorize()
        F = oracle.ilun(A, order) if order else None
        ref = F.vals if order else oracle.ilu0(A)
        assert np.array_equal(M.ilu_values(), ref)
        rs = np.random.RandomState(4)
        for k in range(3):
            v = rs.standard_normal(A.n)
            if k == 2:
                v[::3] = 0.0; v[1::5] = -0.0
            want = oracle.lu_precond(A, F if order else ref, v)
            assert np.array_equal(M.lu_precondition(v).view(np.int64), want.view(np.int64))
        M.set_symmetric_ilu(False); M.factorize()                     # back to the LU branch on the same handle
        oracle.set_cholesky(False)
        want = oracle.lu_precond(A, oracle.ilun(A, order) if order else oracle.ilu0(A), v)
        assert np.array_equal(M.lu_precondition(v), want)
    finally:
        M.close()


@pytest.mark.gpu
def test_gpu_multi_dof_and_unsymmetric_pattern_values(oracle, b200, cholesky):
    """3 dofs per node (node-lane plans must step aside) and a matrix whose VALUES are not symmetric: the factorisation reads the lower
    part only, exactly as the reference."""
    H, _ = _spd(oracle, (5, 4, 4))
    Hs = sp.csr_matrix((H.vals, H.cols - 1, H.rows - 1), shape=(H.n, H.n))
    K = sp.kron(Hs, np.array([[4.0, 1.0, 0.5], [1.0, 3.0, 0.2], [0.5, 0.2, 2.0]])).tocsr()     # SPD, 3 dofs per node, full 3 x 3 blocks
    K.sort_indices()
    rs = np.random.RandomState(2)
    vals = K.data * (1.0 + 1e-3 * rs.standard_normal(K.nnz))                                  # values no longer symmetric
    rows = (K.indptr + 1).astype(np.int32); cols = (K.indices + 1).astype(np.int32)
    diag = np.array([rows[i] + int(np.searchsorted(K.indices[K.indptr[i]:K.indptr[i + 1]], i)) for i in range(K.shape[0])], dtype=np.int32)
    A = oracle.CRS(rows, cols, diag, vals, 3)
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, 3)
        M.set_values(A.vals)
        M.set_symmetric_ilu(True)
        M.factorize()
        ref = oracle.ilu0(A)
        assert np.array_equal(M.ilu_values(), ref)
        v = rs.standard_normal(A.n)
        assert np.array_equal(M.lu_precondition(v), oracle.lu_precond(A, ref, v))
    finally:
        M.close()


@pytest.mark.gpu
def test_gpu_cg_with_the_keyword(oracle, b200, cholesky):
    A, b = _spd(oracle, (24, 24, 24))
    oracle.set_dot_order(3)
    try:
        ref = oracle.itersolve(A, b, method="cg", precond="ilu0", tol=1e-9, maxit=400)
    finally:
        oracle.set_dot_order(0)
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, 1)
        M.set_values(A.vals)
        sif = ("Linear System Iterative Method = CG\nLinear System Max Iterations = 400\nLinear System Convergence Tolerance = 1e-9\n"
               "Linear System Preconditioning = ILU0\nLinear System Symmetric ILU = True\n")
        got = M.itersolver(b, None, sif)
        assert got is not None and got["info"] == ref["info"] == 1 and got["iters"] == ref["iters"]
        assert np.array_equal(got["x"], ref["x"])
    finally:
        M.close()

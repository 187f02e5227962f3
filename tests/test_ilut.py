"""ILUT ('Linear System Preconditioning = ILUT', CRS_ILUT / ComputeILUT, fem/src/CRSMatrix.F90:4144-4340): the factor's pattern is decided
by the values.  CPU: the oracle's restatement against dense LU (tolerance 0 keeps everything: complete LU) and against a dense
threshold-ILU written from the definition.  GPU: pattern, values and both sweeps of the library bit-identical to the restatement; Krylov
counts; the keyword path."""
import numpy as np
import pytest
import scipy.sparse as sp


def _dense(F):
    return sp.csr_matrix((F.vals, F.cols - 1, F.rows - 1), shape=(F.n, F.n)).toarray()


def _dense_ilut(M, tol):
    """Row-by-row IKJ elimination with Elmer's drop rule, on dense rows."""
    n = M.shape[0]
    LU = np.zeros((n, n)); keep = np.zeros((n, n), dtype=bool)
    for i in range(n):
        s = M[i].copy(); flag = M[i] != 0
        for k in range(i):
            if flag[k]:
                if abs(LU[k, k]) > 10 * 2.220446049250313e-16:
                    s[k] = s[k] / LU[k, k]
                up = np.flatnonzero(keep[k, k + 1:]) + k + 1
                flag[up] = True
                s[up] = s[up] - s[k] * LU[k, up]
        norma = np.sqrt(np.sum(np.abs(M[i][M[i] != 0]) ** 2))
        kept = flag & ((np.abs(s) >= tol * norma) | (np.arange(n) == i))
        LU[i, kept] = s[kept]; keep[i] = kept
    return LU, keep


def test_oracle_ilut_tolerance_zero_is_the_complete_lu(oracle):
    A, b = oracle.heat_cube(0, faces=["x0"], dims=(5, 4, 3))
    M = _dense(A)
    F = oracle.ilut(A, 0.0)
    Fm = _dense(F)
    L = np.tril(Fm, -1) + np.eye(A.n); U = np.triu(Fm, 1) + np.diag(1.0 / np.diag(Fm))
    assert np.abs(L @ U - M).max() < 1e-13
    v = np.random.RandomState(1).standard_normal(A.n)
    assert np.abs(M @ oracle.lu_precond(A, F, v) - v).max() < 1e-12


@pytest.mark.parametrize("tol", [1e-3, 1e-2, 0.1])
def test_oracle_ilut_against_the_definition(oracle, tol):
    A, b = oracle.cavity_flow(3)
    M = _dense(A)
    F = oracle.ilut(A, tol)
    LU, keep = _dense_ilut(M, tol)
    Fm = _dense(F)
    pat = sp.csr_matrix((np.ones(F.cols.size), F.cols - 1, F.rows - 1), shape=(A.n, A.n)).toarray() != 0
    assert np.array_equal(pat, keep)
    d = np.diag(LU).copy()
    LU[np.arange(A.n), np.arange(A.n)] = np.where(np.abs(d) < 10 * 2.220446049250313e-16, 1.0, 1.0 / d)
    assert np.abs(Fm - LU).max() <= 1e-12 * np.abs(LU).max()
    assert F.vals.size < oracle.ilut(A, 0.0).vals.size


def test_keyword_is_accepted():
    import elmerfem_b200 as B
    p = B.itersolver_plan("Linear System Max Iterations = 10\nLinear System Iterative Method = GCR\nLinear System Preconditioning = ILUT\n"
                          "Linear System ILUT Tolerance = 1.0e-3\n", 100, 1)
    assert p is not None and p["precond"] == 2


# ---------------------------------------------------------------------------------------------- GPU
def _cases(oracle):
    A1, b1 = oracle.heat_cube(0, faces=["x0"], dims=(8, 7, 6))
    A2, b2 = oracle.cavity_flow(4)
    A3, b3 = oracle.elasticity_beam(6, 3, 3)
    return [(A1, b1, 1), (A2, b2, 4), (A3, b3, 3)]


@pytest.mark.gpu
@pytest.mark.parametrize("tol", [0.0, 1e-4, 1e-2, 0.3])
def test_gpu_factor_pattern_values_and_solve_bit_exact(oracle, b200, tol):
    for A, b, ndeg in _cases(oracle):
        if tol == 0.0 and A.n > 1500:
            continue                                               # complete LU: rows beyond the 1024-entry working row
        F = oracle.ilut(A, tol)
        M = b200.Matrix()
        try:
            M.set_structure(A.rows, A.cols, A.diag, 1, ndeg)
            M.set_values(A.vals)
            M.set_ilut(tol)
            M.factorize()
            rows, cols, diag = M.ilu_structure()
            assert np.array_equal(rows, F.rows) and np.array_equal(cols, F.cols) and np.array_equal(diag, F.diag)
            assert np.array_equal(M.ilu_values().view(np.int64), F.vals.view(np.int64))
            v = np.random.RandomState(3).standard_normal(A.n)
            assert np.array_equal(M.lu_precondition(v), oracle.lu_precond(A, F, v))
            A2 = A.copy(); A2.vals = A.vals * (1.0 + 0.05 * np.sin(np.arange(A.nnz)))      # new values: new pattern
            M.set_values(A2.vals); M.factorize()
            F2 = oracle.ilut(A2, tol)
            rows, cols, diag = M.ilu_structure()
            assert np.array_equal(cols, F2.cols) and np.array_equal(M.ilu_values(), F2.vals)
            M.set_ilut(0.0, on=False); M.factorize()                                         # back to ILU0 on the same handle
            assert np.array_equal(M.ilu_values(), oracle.ilu0(A2))
        finally:
            M.close()


@pytest.mark.gpu
def test_gpu_gcr_with_the_keyword(oracle, b200):
    A, b = oracle.cavity_flow(5)
    F = oracle.ilut(A, 1e-3)
    oracle.set_dot_order(3)
    try:
        ref = oracle.itersolve(A, b, method="gcr", precond="ilu0", ilu=F, tol=1e-9, maxit=300)
    finally:
        oracle.set_dot_order(0)
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, 4)
        M.set_values(A.vals)
        sif = ("Linear System Iterative Method = GCR\nLinear System Max Iterations = 300\nLinear System Convergence Tolerance = 1e-9\n"
               "Linear System Preconditioning = ILUT\nLinear System ILUT Tolerance = 1.0e-3\n")
        got = M.itersolver(b, None, sif)
        assert got is not None and got["info"] == ref["info"] == 1 and got["iters"] == ref["iters"]
        assert np.array_equal(got["x"], ref["x"])
    finally:
        M.close()


@pytest.mark.gpu
def test_gpu_row_overflow_is_an_error_not_a_wrong_factor(oracle, b200):
    A, b = oracle.heat_cube(0, faces=["x0"], dims=(40, 40, 2))           # complete LU of a 41 x 41 x 3 grid: rows of up to ~1700 entries
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, 1)
        M.set_values(A.vals)
        M.set_ilut(0.0)
        with pytest.raises(b200.B200Error):
            M.factorize()
    finally:
        M.close()

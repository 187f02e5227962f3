"""GPU leg of the matrix-structure producer: a system whose Rows/Cols/Diag come from
b200_node_graph / b200_optimize_bandwidth / b200_initialize_structure (Elmer's default `Optimize Bandwidth`
numbering, accepted on a beam) goes through the C ABI like any other matrix; the ordering changes the ILU0 factor
and the dependency levels, so factor, triangular solve and Krylov parity are checked again on it."""
import numpy as np
import pytest

import structure_case as sc

pytestmark = pytest.mark.gpu

TOL = 1e-8


def test_elmer_ordered_beam_parity(oracle, b200):
    S = sc.beam_heat_in_elmer_order()
    assert S["accepted"]
    A, b = S["A"].copy(), S["b"].copy()
    oracle.scale_system(A, b, np.zeros(A.n))
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg)
        M.set_values(A.vals)
        r, c, d = M.structure()
        assert np.array_equal(r, A.rows) and np.array_equal(c, A.cols) and np.array_equal(d, A.diag)
        u = np.random.RandomState(5).standard_normal(A.n)
        assert np.array_equal(M.matvec(u), oracle.matvec(A, u))
        M.factorize()
        ilu = oracle.ilu0(A)
        assert np.array_equal(M.ilu_values(), ilu)
        assert np.array_equal(M.lu_precondition(u), oracle.lu_precond(A, ilu, u))
        ref = oracle.itersolve(A, b, method="bicgstab", precond="ilu0", tol=TOL, maxit=500)
        got = M.solve(b, method="bicgstab", precond="ilu0", tol=TOL, maxit=500)
        assert ref["info"] == 1 and got["info"] == 1
        assert abs(got["iters"] - ref["iters"]) <= max(1, int(np.ceil(0.02 * ref["iters"])))
        assert np.linalg.norm(got["x"] - ref["x"]) <= 10 * TOL * np.linalg.norm(ref["x"])
    finally:
        M.close()

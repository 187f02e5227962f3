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


# ---- `Linear System Robust` on the device (IterativeMethods.F90:1110-1139, 1862-1898): same stop, same iterate ----------
ROBUST_CASES = [("bicgstabl", dict(bicgstabl_l=2), "none"), ("bicgstabl", dict(bicgstabl_l=2), "ilu0"),
                ("idrs", dict(idrs_s=4), "none")]


@pytest.mark.parametrize("method,kw,pc", ROBUST_CASES)
def test_robust_mode_parity(oracle, b200, method, kw, pc):
    A, b = oracle.cavity_flow(6)
    A = A.copy()
    oracle.scale_system(A, b, np.zeros(A.n))
    P = oracle.shadow_space(A.n, 4) if method == "idrs" else None
    opts = dict(tol=1e-10, maxit=400, robust=True, robust_tol=1e-4, robust_limit=1e-2, robust_max_bad=0, **kw)
    ref = oracle.itersolve(A, b, method=method, precond=pc, P=P, **opts)
    plain = oracle.itersolve(A, b, method=method, precond=pc, P=P, tol=1e-10, maxit=400, **kw)
    assert ref["iters"] < plain["iters"]                    # the safeguard does fire in this case
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg)
        M.set_values(A.vals)
        got = M.solve(b, method=method, precond=pc, P=P, **opts)
        assert got["info"] == ref["info"] == 1
        close = lambda a, c: abs(a - c) <= max(1, int(np.ceil(0.02 * max(a, c))))      # the 2 % bar of BASELINE.json
        assert close(got["iters"], ref["iters"]), (got["iters"], ref["iters"])
        assert got["iters"] < plain["iters"]
        true = lambda x: float(np.linalg.norm(oracle.matvec(A, x) - b) / np.linalg.norm(b))
        assert true(got["x"]) < 1e-4 and true(got["x"]) <= got["residual"] * (1 + 1e-6)    # the best iterate came back
        if got["iters"] == ref["iters"]:                     # same stop => same iterate (an UNconverged one: rounding
            assert np.linalg.norm(got["x"] - ref["x"]) <= 1e-6 * np.linalg.norm(ref["x"])   # differences are not damped)
        # keyword path: no longer declined
        sif = ("Linear System Iterative Method = %s\nLinear System Preconditioning = %s\nLinear System Max Iterations = 400\n"
               "Linear System Convergence Tolerance = 1e-10\nLinear System Robust = True\nLinear System Robust Tolerance = 1e-4\n"
               "Linear System Robust Limit = 1e-2\nLinear System Robust Max Iterations = 0\n" % (method, pc))
        if method == "bicgstabl":
            sif += "BiCGstabl polynomial degree = %d\n" % kw["bicgstabl_l"]
            out = M.itersolver(b, None, sif)
            assert out is not None and out["info"] == 1 and out["iters"] == got["iters"]
            assert np.array_equal(out["x"], got["x"])        # same device path, same bits
    finally:
        M.close()

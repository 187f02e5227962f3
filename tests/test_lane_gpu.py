"""Lane-tile triangular solve (csrc/lane.cu, B200_TRI_MODE=4) on the GPU through the C ABI: bit-identical to the oracle's CRS_LUSolve on
grids that exercise single tiles, many tiles, partial strips, several tiles per warp and one and two planes per lane; BiCGStab + ILU0 on top
of it takes the oracle's iteration count; anything that is not the 27-point grid stencil falls back to the level kernel.  CPU
counterpart: tests/test_lane_plan.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case(oracle, dims):
    A, b = oracle.heat_cube(0, faces=["x0"], dims=dims)
    A = A.copy()
    oracle.scale_system(A, b, np.zeros(A.n))
    return A, b


@pytest.mark.parametrize("tc", [1, 2])
@pytest.mark.parametrize("dims", [(6, 6, 6), (9, 4, 5), (40, 40, 3), (33, 70, 12), (50, 20, 40)])
def test_lane_mode_bit_exact(oracle, b200, monkeypatch, dims, tc):
    monkeypatch.setenv("B200_TRI_MODE", "4")
    monkeypatch.setenv("B200_LANE_TC", str(tc))
    A, b = _case(oracle, dims)
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, 1)
        M.set_values(A.vals)
        M.factorize()
        assert M.stats()["tri_mode"] == 4
        ilu = oracle.ilu0(A)
        assert np.array_equal(M.ilu_values(), ilu)
        for seed in (1, 2, 3):
            v = np.random.RandomState(seed).standard_normal(A.n)
            if seed == 3:
                v[::3] = 0.0; v[1::7] = -0.0                      # exact and signed zeros keep their bits
            got, ref = M.lu_precondition(v), oracle.lu_precond(A, ilu, v)
            assert np.array_equal(got.view(np.int64), ref.view(np.int64))
    finally:
        M.close()


@pytest.mark.parametrize("warps,depth", [(1, 1), (3, 3), (8, 7)])
def test_lane_mode_few_warps_many_rounds(oracle, b200, monkeypatch, warps, depth):
    """More tiles than co-resident warps (static round-robin over start levels) and every request lead."""
    monkeypatch.setenv("B200_TRI_MODE", "4")
    monkeypatch.setenv("B200_LANE_TC", "1")
    monkeypatch.setenv("B200_LANE_WARPS", str(warps))
    monkeypatch.setenv("B200_LANE_E", str(depth))
    A, b = _case(oracle, (20, 70, 90))
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, 1)
        M.set_values(A.vals)
        M.factorize()
        ilu = oracle.ilu0(A)
        v = np.random.RandomState(5).standard_normal(A.n)
        assert np.array_equal(M.lu_precondition(v), oracle.lu_precond(A, ilu, v))
    finally:
        M.close()


def test_lane_mode_krylov_and_refactorisation(oracle, b200, monkeypatch):
    monkeypatch.setenv("B200_TRI_MODE", "4")
    A, b = _case(oracle, (30, 30, 30))
    oracle.set_dot_order(3)
    try:
        ref = oracle.itersolve(A, b, method="bicgstab", precond="ilu0", tol=1e-8, maxit=500)
    finally:
        oracle.set_dot_order(0)
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, 1)
        M.set_values(A.vals)
        got = M.solve(b, method="bicgstab", precond="ilu0", tol=1e-8, maxit=500)
        assert got["info"] == ref["info"] == 1 and got["iters"] == ref["iters"]
        assert np.array_equal(got["x"], ref["x"])
        A2 = A.copy(); A2.vals = A.vals * (1.0 + 0.01 * np.cos(np.arange(A.nnz)))   # new values, same structure: streams are refilled
        M.set_values(A2.vals); M.factorize()
        v = np.random.RandomState(8).standard_normal(A.n)
        assert np.array_equal(M.lu_precondition(v), oracle.lu_precond(A2, oracle.ilu0(A2), v))
    finally:
        M.close()


def test_lane_mode_falls_back_when_the_structure_is_not_a_grid_stencil(oracle, b200, monkeypatch):
    monkeypatch.setenv("B200_TRI_MODE", "4")
    A, b = oracle.elasticity_beam(4, 3, 3)
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, 3)
        M.set_values(A.vals)
        M.factorize()
        ilu = oracle.ilu0(A)
        v = np.random.RandomState(4).standard_normal(A.n)
        assert np.array_equal(M.lu_precondition(v), oracle.lu_precond(A, ilu, v))      # level kernel
    finally:
        M.close()

"""Opt-in skewed-lane triangular solve (csrc/skew.cu, B200_TRI_MODE=2) on the GPU.  NOT part of `-m gpu`: the kernel was written at
the end of round 1 without GPU budget left; its schedule and stream layout are verified on the CPU (tests/test_skew_plan.py).  First run:
    gpurun --timeout 300 -- 'timeout 120 python -m pytest tests/test_experimental_skew_gpu.py -m gpu_experimental -x -q'
The kernel spins on L2 sentinels with a bounded spin count (HUTI_HALTED / an error on timeout), so a wrong dependency shows up as a
failed assertion, not as a hang -- still, run it under `timeout`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu_experimental


@pytest.mark.parametrize("dims", [(6, 6, 6), (9, 4, 5), (40, 40, 3), (33, 70, 2)])
def test_skew_mode_bit_exact(oracle, b200, monkeypatch, dims):
    monkeypatch.setenv("B200_TRI_MODE", "2")
    monkeypatch.setenv("B200_SKEW_DEBUG", "1")
    A, b = oracle.heat_cube(0, faces=["x0"], dims=dims)
    A = A.copy()
    oracle.scale_system(A, b, np.zeros(A.n))
    M = b200.Matrix()
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, 1)
        M.set_values(A.vals)
        M.factorize()
        ilu = oracle.ilu0(A)
        assert np.array_equal(M.ilu_values(), ilu)
        for seed in (1, 2):
            v = np.random.RandomState(seed).standard_normal(A.n)
            assert np.array_equal(M.lu_precondition(v), oracle.lu_precond(A, ilu, v))
        ref = oracle.itersolve(A, b, method="bicgstab", precond="ilu0", tol=1e-8, maxit=500)
        got = M.solve(b, method="bicgstab", precond="ilu0", tol=1e-8, maxit=500)
        assert got["info"] == ref["info"] == 1 and abs(got["iters"] - ref["iters"]) <= 1
        assert np.linalg.norm(got["x"] - ref["x"]) <= 1e-7 * np.linalg.norm(ref["x"])
    finally:
        M.close()


def test_skew_mode_falls_back_when_the_structure_is_not_a_grid_stencil(oracle, b200, monkeypatch):
    monkeypatch.setenv("B200_TRI_MODE", "2")
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

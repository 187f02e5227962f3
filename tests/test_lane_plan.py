"""The lane-tile triangular solve (csrc/lane.cu) on the CPU: tests/lane_harness.cpp executes the kernel's schedule through the kernel's own
geometry / layout / shuffle-routing / row-arithmetic header (csrc/lanegeom.h) on real ILU0 factors; the result must be bit-identical to
the oracle's CRS_LUSolve, no replayed value may come from a tile that has not run yet (the tile order would deadlock), and matrices
without the grid stencil must be refused (the level kernel then stays)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lane") / "lane_harness.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "lane_harness.cpp")])
    L = C.CDLL(so)
    L.lane_emulate.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp, _dp, _ip, C.c_int]
    return L


def _run(harness, A, ilu, v, TC):
    x = np.zeros(A.n); geom = np.zeros(5, dtype=np.int32)
    rc = harness.lane_emulate(A.n, A.rows - 1, A.cols - 1, A.diag - 1, np.ascontiguousarray(ilu), np.ascontiguousarray(v), x, geom, TC)
    return rc, x, geom


@pytest.mark.parametrize("dims", [(6, 6, 6), (9, 4, 5), (3, 40, 4), (12, 12, 1), (35, 3, 3), (4, 33, 19), (2, 65, 9), (2, 2, 2)])
def test_bit_identical_to_crs_lusolve(oracle, harness, dims):
    A, b = oracle.heat_cube(0, faces=["x0"], dims=dims)
    A = A.copy()
    oracle.scale_system(A, b, np.zeros(A.n))
    ilu = oracle.ilu0(A)
    v = np.random.RandomState(3).standard_normal(A.n)
    ref = oracle.lu_precond(A, ilu, v)
    for TC in (1, 2, 3, 4):
        rc, x, geom = _run(harness, A, ilu, v, TC)
        assert rc == 0, (rc, TC)
        assert tuple(geom[:3]) == (dims[0] + 1, dims[1] + 1, dims[2] + 1)
        assert np.array_equal(x, ref), TC


def test_signed_zeros_and_exact_zero_rows(oracle, harness):
    """Pad entries must be (+0) x (+0): a right-hand side with exact (signed) zeros keeps every bit."""
    A, b = oracle.heat_cube(0, faces=["x0"], dims=(5, 6, 4))
    A = A.copy()
    oracle.scale_system(A, b, np.zeros(A.n))
    ilu = oracle.ilu0(A)
    v = np.random.RandomState(9).standard_normal(A.n)
    v[::3] = 0.0; v[1::7] = -0.0
    rc, x, _ = _run(harness, A, ilu, v, 2)
    assert rc == 0
    ref = oracle.lu_precond(A, ilu, v)
    assert np.array_equal(x.view(np.int64), ref.view(np.int64))


def test_other_structures_are_refused(oracle, harness):
    A, b = oracle.elasticity_beam(3, 3, 3)                      # 3 dofs per node: not the scalar stencil
    rc, _, _ = _run(harness, A, oracle.ilu0(A), np.ones(A.n), 2)
    assert rc == 1

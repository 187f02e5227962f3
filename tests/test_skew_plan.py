"""The skewed-lane triangular solve (csrc/skew.cu) on the CPU: tests/skew_harness.cpp executes the kernel's schedule through the kernel's
own geometry / detection / layout / operand-routing header (csrc/skewgeom.h) on real ILU0 factors; the result must be bit-identical to
the oracle's CRS_LUSolve, and matrices without the grid stencil must be refused (the level kernel then stays)."""
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
    so = str(tmp_path_factory.mktemp("skew") / "skew_harness.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "skew_harness.cpp")])
    L = C.CDLL(so)
    L.skew_emulate.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp, _dp, _ip, C.c_int]
    L.skew_emulate_concurrent.argtypes = [C.c_int, _ip, _ip, _ip, _dp, _dp, _dp, C.c_int, C.c_uint, C.c_int]
    return L


def _run(harness, A, ilu, v, E=3):
    x = np.zeros(A.n); geom = np.zeros(5, dtype=np.int32)
    rc = harness.skew_emulate(A.n, A.rows - 1, A.cols - 1, A.diag - 1, np.ascontiguousarray(ilu), np.ascontiguousarray(v), x, geom, E)
    return rc, x, geom


@pytest.mark.parametrize("dims", [(6, 6, 6), (9, 4, 5), (3, 40, 4), (12, 12, 1), (35, 3, 3)])
def test_bit_identical_to_crs_lusolve(oracle, harness, dims):
    A, b = oracle.heat_cube(0, faces=["x0"], dims=dims)
    A = A.copy()
    oracle.scale_system(A, b, np.zeros(A.n))
    ilu = oracle.ilu0(A)
    v = np.random.RandomState(3).standard_normal(A.n)
    for E in (3, 1, 7):                                          # request lead of the kernel configurations
        rc, x, geom = _run(harness, A, ilu, v, E)
        assert rc == 0
        assert tuple(geom[:3]) == (dims[0] + 1, dims[1] + 1, dims[2] + 1)
        assert geom[3] <= 29 and geom[3] * geom[4] >= geom[1]
        assert np.array_equal(x, oracle.lu_precond(A, ilu, v))


def test_other_structures_are_refused(oracle, harness):
    A, b = oracle.elasticity_beam(3, 3, 3)                      # 3 dofs per node: not the scalar stencil
    rc, _, _ = _run(harness, A, oracle.ilu0(A), np.ones(A.n))
    assert rc == 1
    import scipy.sparse as sp
    H, _ = oracle.heat_cube(5, faces=["x0"])
    p = np.random.RandomState(0).permutation(H.n)
    S = H.to_scipy()[p][:, p]                                    # same graph, scrambled numbering
    Hs = oracle.CRS.from_scipy(sp.csr_matrix(S))
    rc, _, _ = _run(harness, Hs, oracle.ilu0(Hs), np.ones(Hs.n))
    assert rc == 1


@pytest.mark.parametrize("dims,nw", [((6, 6, 6), 1), ((6, 6, 6), 2), ((5, 70, 4), 2), ((5, 70, 4), 3), ((5, 70, 4), 7), ((9, 40, 5), 5),
                                     ((9, 40, 5), 64), ((4, 4, 12), 4)])
def test_no_deadlock_for_any_warp_count_or_interleaving(oracle, harness, dims, nw):
    """The kernel's task assignment (warp w takes tasks w, w+NW, ... in increasing order) with the sentinel protocol: emulated warps are
    interleaved pseudo-randomly and may only advance when their L2 operands exist.  Fewer warps than strips per plane, one warp, more
    warps than tasks: every run must finish (no deadlock) with the bit-identical result."""
    A, b = oracle.heat_cube(0, faces=["x0"], dims=dims)
    A = A.copy()
    oracle.scale_system(A, b, np.zeros(A.n))
    ilu = oracle.ilu0(A)
    v = np.random.RandomState(5).standard_normal(A.n)
    ref = oracle.lu_precond(A, ilu, v)
    for seed in (1, 2, 3):
        rc = harness.skew_emulate_concurrent(A.n, A.rows - 1, A.cols - 1, A.diag - 1, np.ascontiguousarray(ilu), np.ascontiguousarray(v),
                                             np.ascontiguousarray(ref), nw, seed, 3)
        assert rc == 0, (rc, seed)

"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): integer work bit-exact; iteration counts within 2 %; converged
solution's relative L2 difference <= 10 x Linear System Convergence Tolerance.  The row-wise kernels
(SpMV, ILU0 factor, triangular solves, Jacobi) reproduce the reference's operation order, so for
them the bar used here is bit-exact.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-8


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def iters_close(a, b):
    return abs(a - b) <= max(1, int(np.ceil(0.02 * max(a, b))))


@pytest.fixture(scope="module")
def heat(oracle):
    A, b = oracle.heat_cube(24, faces=["x0"], source=1.0)
    A = A.copy()
    x = np.zeros(A.n)
    oracle.scale_system(A, b, x)          # Elmer's default Linear System Scaling
    return A, b


@pytest.fixture(scope="module")
def heat_gpu(b200, heat):
    A, b = heat
    M = b200.Matrix()
    M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg)
    M.set_values(A.vals)
    yield M
    M.close()


def test_structure_mirror_bit_exact(heat, heat_gpu):
    A, _ = heat
    r, c, d = heat_gpu.structure()
    assert np.array_equal(r, A.rows) and np.array_equal(c, A.cols) and np.array_equal(d, A.diag)


def test_spmv_bit_exact(oracle, heat, heat_gpu):
    A, _ = heat
    rng = np.random.RandomState(1)
    for _ in range(3):
        u = rng.standard_normal(A.n)
        assert np.array_equal(heat_gpu.matvec(u), oracle.matvec(A, u))


def test_diag_precondition_bit_exact(oracle, b200):
    A, b = oracle.heat_cube(10, faces=["x0"])      # unscaled: diagonal != 1
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag); M.set_values(A.vals)
    v = np.random.RandomState(2).standard_normal(A.n)
    assert np.array_equal(M.diag_precondition(v), oracle.diag_precond(A, v))
    M.close()


def test_ilu0_factor_bit_exact(oracle, heat, heat_gpu):
    A, _ = heat
    heat_gpu.factorize()
    assert np.array_equal(heat_gpu.ilu_values(), oracle.ilu0(A))


def test_lu_solve_bit_exact(oracle, heat, heat_gpu):
    A, _ = heat
    ilu = oracle.ilu0(A)
    v = np.random.RandomState(3).standard_normal(A.n)
    assert np.array_equal(heat_gpu.lu_precondition(v), oracle.lu_precond(A, ilu, v))


def test_levels_of_structured_grid(heat_gpu):
    # forward level of node (a,b,c) on an x-fastest hex8 grid is a + 2b + 4c
    n1 = 25
    lv = heat_gpu.levels(per_row=True)
    a, b, c = np.meshgrid(np.arange(n1), np.arange(n1), np.arange(n1), indexing="ij")
    expect = (a + 2 * b + 4 * c).transpose(2, 1, 0).reshape(-1)
    assert np.array_equal(lv["level"], expect)
    assert lv["forward"] == 7 * (n1 - 1) + 1


def test_dot_and_norm(oracle, heat_gpu):
    rng = np.random.RandomState(4)
    x = rng.standard_normal(heat_gpu.n); y = rng.standard_normal(heat_gpu.n)
    assert abs(heat_gpu.dot(x, y) - oracle.ddot(x, y)) <= 1e-12 * np.linalg.norm(x) * np.linalg.norm(y)
    assert abs(heat_gpu.nrm2(x) - oracle.dnrm2(x)) <= 1e-13 * oracle.dnrm2(x)


@pytest.mark.parametrize("method", ["cg", "bicgstab", "bicgstabl", "gcr", "idrs"])
@pytest.mark.parametrize("precond", ["none", "diagonal", "ilu0"])
def test_krylov_parity_heat(oracle, heat, heat_gpu, method, precond):
    A, b = heat
    kw = dict(tol=TOL, maxit=2000, bicgstabl_l=4)
    P = oracle.shadow_space(A.n, 4) if method == "idrs" else None
    ref = oracle.itersolve(A, b, method=method, precond=precond, P=P, **kw)
    got = heat_gpu.solve(b, method=method, precond=precond, P=P, **kw)
    assert ref["info"] == 1 and got["info"] == 1, (ref["info"], got["info"])
    assert iters_close(got["iters"], ref["iters"]), (got["iters"], ref["iters"])
    assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
    # true residual of the GPU answer
    r = oracle.matvec(A, got["x"]) - b
    assert np.linalg.norm(r) / np.linalg.norm(b) <= 20 * TOL


def test_call_counts_match_reference(oracle, heat, heat_gpu):
    A, b = heat
    ref = oracle.itersolve(A, b, method="bicgstab", precond="ilu0", tol=TOL, maxit=500)
    got = heat_gpu.solve(b, method="bicgstab", precond="ilu0", tol=TOL, maxit=500)
    assert got["iters"] == ref["iters"], (got["iters"], ref["iters"])
    assert got["stats"]["matvec"] == ref["counts"]["matvec"]
    assert got["stats"]["pcond"] == ref["counts"]["pcond"]


def test_maxiter_and_info_codes(oracle, heat, heat_gpu):
    A, b = heat
    P = oracle.shadow_space(A.n, 4)
    for method in ["cg", "bicgstab", "bicgstabl", "gcr", "idrs"]:
        ref = oracle.itersolve(A, b, method=method, precond="none", tol=1e-14, maxit=3, bicgstabl_l=2, P=P)
        got = heat_gpu.solve(b, method=method, precond="none", tol=1e-14, maxit=3, bicgstabl_l=2, P=P)
        assert got["info"] == ref["info"] == 2, (method, got["info"], ref["info"])
        assert got["iters"] == ref["iters"], (method, got["iters"], ref["iters"])
        assert rel_l2(got["x"], ref["x"]) < 1e-10


def test_elasticity_ndeg3(oracle, b200):
    A, b = oracle.elasticity_beam(12, 4, 4, lx=3.0)
    A = A.copy(); x = np.zeros(A.n); oracle.scale_system(A, b, x)
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 3); M.set_values(A.vals)
    u = np.random.RandomState(5).standard_normal(A.n)
    assert np.array_equal(M.matvec(u), oracle.matvec(A, u))          # the 3-accumulator variant, bit for bit
    M.factorize()
    assert np.array_equal(M.ilu_values(), oracle.ilu0(A))
    ref = oracle.itersolve(A, b, method="bicgstabl", precond="ilu0", tol=TOL, maxit=500, bicgstabl_l=4)
    got = M.solve(b, method="bicgstabl", precond="ilu0", tol=TOL, maxit=500, bicgstabl_l=4)
    assert got["info"] == ref["info"] == 1
    assert iters_close(got["iters"], ref["iters"])
    assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
    M.close()


def test_nonsymmetric_cavity(oracle, b200):
    A, b = oracle.cavity_flow(6)
    A = A.copy(); x = np.zeros(A.n); oracle.scale_system(A, b, x)
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 4); M.set_values(A.vals)
    u = np.random.RandomState(6).standard_normal(A.n)
    assert np.array_equal(M.matvec(u), oracle.matvec(A, u))
    for method in ["gcr", "idrs", "bicgstabl"]:
        P = oracle.shadow_space(A.n, 4) if method == "idrs" else None
        ref = oracle.itersolve(A, b, method=method, precond="ilu0", tol=TOL, maxit=1000, bicgstabl_l=4, P=P)
        got = M.solve(b, method=method, precond="ilu0", tol=TOL, maxit=1000, bicgstabl_l=4, P=P)
        assert got["info"] == ref["info"], (method, got["info"], ref["info"])
        if ref["info"] == 1:
            assert iters_close(got["iters"], ref["iters"]), (method, got["iters"], ref["iters"])
            assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
    M.close()


def test_testmat_known_answer(oracle, b200):
    """fhutiter/examples/ex1/testmat(.out): committed copy under tests/golden/."""
    import os
    import scipy.sparse as sp
    g = os.path.join(os.path.dirname(__file__), "golden")
    d = np.loadtxt(os.path.join(g, "huti_ex1_testmat.txt"))
    xref = np.loadtxt(os.path.join(g, "huti_ex1_testmat_out.txt"))[:, 1]
    Ms = sp.coo_matrix((d[:, 2], (d[:, 0].astype(int) - 1, d[:, 1].astype(int) - 1)), shape=(100, 100)).tocsr()
    A = oracle.CRS.from_scipy(Ms)
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag); M.set_values(A.vals)
    for method in ["bicgstabl", "gcr", "idrs"]:
        got = M.solve(np.ones(100), method=method, precond="none", tol=1e-10, maxit=500, bicgstabl_l=4)
        assert got["info"] == 1
        assert np.abs(got["x"] - xref).max() < 1e-6
    M.close()


def test_spmv_hook_matches_load_c_signature(oracle, b200, heat):
    A, _ = heat
    slot = C.c_void_p(None)
    u = np.random.RandomState(7).standard_normal(A.n)
    v = b200.spmv_hook(slot, A.rows, A.cols, A.vals, u)
    assert np.array_equal(v, oracle.matvec(A, u))
    vals2 = A.vals * 2.0                       # the hook is never told that Values changed
    v2 = b200.spmv_hook(slot, A.rows, A.cols, vals2, u)
    assert np.array_equal(v2, 2.0 * v)
    # changes IN PLACE (same pointer) that a floating-point checksum can miss: a tiny entry next to large ones, a swap 1024 entries apart
    vals3 = vals2.copy()
    b200.spmv_hook(slot, A.rows, A.cols, vals3, u)
    vals3[5] = np.nextafter(vals3[5], np.inf)                        # one ulp
    A3 = A.copy(); A3.vals = vals3
    assert np.array_equal(b200.spmv_hook(slot, A.rows, A.cols, vals3, u), oracle.matvec(A3, u))
    vals3[[7, 7 + 1024]] = vals3[[7 + 1024, 7]]
    assert np.array_equal(b200.spmv_hook(slot, A.rows, A.cols, vals3, u), oracle.matvec(A3, u))
    b200.lib().b200_destroy(C.byref(slot))


def test_itersolver_keywords(oracle, heat, heat_gpu):
    A, b = heat
    sif = """
      Linear System Solver = Iterative
      Linear System Iterative Method = BiCGStab
      Linear System Preconditioning = ILU0
      Linear System Max Iterations = 500
      Linear System Convergence Tolerance = 1.0e-8
      Linear System Residual Output = 0
    """
    got = heat_gpu.itersolver(b, None, sif, solve_count=0)
    ref = oracle.itersolve(A, b, method="bicgstab", precond="ilu0", tol=1e-8, maxit=500)
    assert got is not None and got["info"] == 1 and got["solve_count"] == 1
    assert iters_close(got["iters"], ref["iters"])
    assert rel_l2(got["x"], ref["x"]) <= 1e-7
    declined = heat_gpu.itersolver(b, None, sif.replace("ILU0", "Multigrid"), 0)
    assert declined is None
    assert heat_gpu.itersolver(b, None, sif + "\n      Linear System Complex = True\n", 0) is None


def test_empty_and_tiny_systems(oracle, b200):
    import scipy.sparse as sp
    for n in [1, 2, 33]:
        Ms = sp.diags([np.full(n, 4.0)] + ([np.full(n - 1, -1.0)] * 2 if n > 1 else []), [0] + ([-1, 1] if n > 1 else [])).tocsr()
        A = oracle.CRS.from_scipy(Ms)
        M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag); M.set_values(A.vals)
        b = np.arange(1, n + 1, dtype=float)
        for method in ["cg", "bicgstab"]:
            ref = oracle.itersolve(A, b, method=method, precond="ilu0", tol=1e-10, maxit=50)
            got = M.solve(b, method=method, precond="ilu0", tol=1e-10, maxit=50)
            assert got["info"] == ref["info"]
            assert rel_l2(got["x"], ref["x"]) < 1e-9
        M.close()


@pytest.mark.parametrize("mode", ["0", "-2"])
def test_triangular_solves_bit_exact_on_other_structures(oracle, b200, heat, mode, monkeypatch):
    """The level kernel (B200_TRI_MODE=0) and the default (-2: level / wave / lane kernels timed, the fastest kept, anything that is not a
    grid stencil stays with the level kernel) must give the reference's CRS_LUSolve bit for bit: heat (13 lower entries), elasticity
    (rows wider than the 16-operand register chunk), a nonsymmetric 4-dof pattern, and chains without any parallelism (tridiagonal)."""
    import scipy.sparse as sp
    monkeypatch.setenv("B200_TRI_MODE", mode)
    cases = [heat[0]]
    A2, b2 = oracle.elasticity_beam(10, 4, 4, lx=2.5); cases.append(A2)
    A3, b3 = oracle.cavity_flow(5); cases.append(A3)
    for n in [1, 2, 33, 700]:
        Ms = sp.diags([np.full(n, 4.0)] + ([np.full(n - 1, -1.0)] * 2 if n > 1 else []), [0] + ([-1, 1] if n > 1 else [])).tocsr()
        cases.append(oracle.CRS.from_scipy(Ms))
    for A in cases:
        M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg); M.set_values(A.vals)
        M.factorize()
        ilu = oracle.ilu0(A)
        for seed in (7, 8):
            v = np.random.RandomState(seed).standard_normal(A.n)
            assert np.array_equal(M.lu_precondition(v), oracle.lu_precond(A, ilu, v)), (A.n, A.ndeg)
        M.close()
    A, b = heat
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg); M.set_values(A.vals)
    ref = oracle.itersolve(A, b, method="bicgstab", precond="ilu0", tol=TOL, maxit=500)
    got = M.solve(b, method="bicgstab", precond="ilu0", tol=TOL, maxit=500)
    assert got["info"] == ref["info"] == 1 and got["iters"] == ref["iters"]
    assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
    M.close()


def test_config1_full_size_cg_jacobi(oracle, b200):
    """BASELINE configs[0] at its full size (heat 100^3 hex8, 1,030,301 dofs, CG + Jacobi, tol 1e-8): iteration count and
    solution against the oracle (OpenMP SpMV, a few seconds on the host)."""
    from elmerfem_b200 import synth
    A, b = synth.workload("heat", 100)
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 1); M.set_values(A.vals)
    got = M.solve(b, method="cg", precond="diagonal", tol=TOL, maxit=2000)
    ref = oracle.itersolve(oracle.CRS(A.rows, A.cols, A.diag, A.vals, 1), b, method="cg", precond="diagonal", tol=TOL, maxit=2000)
    assert got["info"] == ref["info"] == 1
    assert iters_close(got["iters"], ref["iters"]), (got["iters"], ref["iters"])
    assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
    u = np.random.RandomState(11).standard_normal(A.n)
    assert np.array_equal(M.matvec(u), oracle.matvec(oracle.CRS(A.rows, A.cols, A.diag, A.vals, 1), u))
    M.close()


def test_config2_full_size_properties(b200):
    """BASELINE configs[1] at its full size (heat 200^3, 8,120,601 dofs, BiCGStab + ILU0): size-independent properties.
    SpMV is linear and matches scipy on the same CRS; L U (M^-1 v) reproduces v; the solve converges and the true
    residual recomputed from the answer meets the tolerance.  Iteration count: EQUAL to the oracle's count for this system when the
    oracle sums its dot products in the device's order (85; `oracle.set_dot_order(3)`).  With the reference's strictly sequential
    ddot the oracle takes 92, with eight interleaved partial sums 90, pairwise 88 (all four recorded in
    tests/golden/c2_device_order.json): the count of BiCGStab on this system moves with the summation order alone, which
    tests/test_gpu_bitwise.py proves by demanding bit-identical solutions."""
    from elmerfem_b200 import synth
    A, b = synth.workload("heat", 200)
    S = A.to_scipy()
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 1); M.set_values(A.vals)
    rs = np.random.RandomState(12)
    u, w = rs.standard_normal(A.n), rs.standard_normal(A.n)
    yu, yw, yuw = M.matvec(u), M.matvec(w), M.matvec(2.0 * u - 0.5 * w)
    ref = S @ u
    assert np.abs(yu - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.abs(yuw - (2.0 * yu - 0.5 * yw)).max() <= 1e-12 * np.abs(yuw).max()
    M.factorize()
    lu = M.ilu_values()
    z = M.lu_precondition(u)
    # rebuild L (unit diagonal) and U (inverse diagonal stored) from ILUValues and check L U z = u
    import scipy.sparse as sp
    rows0, cols0 = np.repeat(np.arange(A.n), np.diff(A.rows)), A.cols - 1
    low, upp = cols0 < rows0, cols0 > rows0
    d = 1.0 / lu[A.diag - 1]
    L = sp.csr_matrix((lu[low], (rows0[low], cols0[low])), shape=S.shape) + sp.identity(A.n, format="csr")
    U = sp.csr_matrix((lu[upp], (rows0[upp], cols0[upp])), shape=S.shape) + sp.diags(d, format="csr")
    back = L @ (U @ z)
    assert np.abs(back - u).max() <= 1e-10 * np.abs(u).max()
    got = M.solve(b, method="bicgstab", precond="ilu0", tol=TOL, maxit=2000)
    # the oracle's count for this very system under the device's summation order (tests/golden/c2_device_order.json, written on the CPU by
    # tests/studies/c2_device_order_golden.py); tests/test_gpu_bitwise.py additionally demands the same bits of the solution
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_device_order.json")))
    assert got["info"] == 1 and got["iters"] == gold["order_3"]["iters"], (got["iters"], gold["order_3"]["iters"])
    r = S @ got["x"] - b
    assert np.linalg.norm(r) / np.linalg.norm(b) <= TOL
    M.close()


@pytest.mark.parametrize("order", [1, 2])
def test_ilun_bit_exact(oracle, b200, heat, order):
    """ILU(n > 0), CRS_IncompleteLU(A, n) (CRSMatrix.F90:3445-3795): the fill pattern (integer work) and ILUValues are
    bit-identical to the oracle, so are the triangular solves on that pattern (both kernels), and the Krylov methods
    preconditioned with it take the oracle's iteration counts.  Also through the SIF keyword path."""
    cases = [heat[0]]
    A2, _ = oracle.elasticity_beam(6, 3, 3, lx=2.0); cases.append(A2)
    A3, _ = oracle.cavity_flow(4); cases.append(A3)
    for A in cases:
        F = oracle.ilun(A, order)
        M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg); M.set_values(A.vals)
        M.set_ilu_order(order)
        r, c, d = M.ilu_structure()
        assert np.array_equal(r, F.rows) and np.array_equal(c, F.cols) and np.array_equal(d, F.diag)
        M.factorize()
        assert np.array_equal(M.ilu_values(), F.vals), (A.n, A.ndeg)
        v = np.random.RandomState(21).standard_normal(A.n)
        assert np.array_equal(M.lu_precondition(v), oracle.lu_precond(A, F, v))
        M.set_ilu_order(0); M.factorize()                       # back to ILU0 on the same handle
        assert np.array_equal(M.ilu_values(), oracle.ilu0(A))
        M.close()
    A, b = heat
    F = oracle.ilun(A, order)
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg); M.set_values(A.vals)
    M.set_ilu_order(order)
    for method in ["cg", "bicgstab", "gcr"]:
        ref = oracle.itersolve(A, b, method=method, precond="ilu%d" % order, ilu=F, tol=TOL, maxit=500)
        got = M.solve(b, method=method, precond="ilu0", tol=TOL, maxit=500)     # precond code 2 = the current ILU order
        assert got["info"] == ref["info"] == 1 and got["iters"] == ref["iters"], (method, got["iters"], ref["iters"])
        assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
    M.close()
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg); M.set_values(A.vals)
    sif = """
      Linear System Solver = Iterative
      Linear System Iterative Method = BiCGStab
      Linear System Preconditioning = ILU%d
      Linear System Max Iterations = 500
      Linear System Convergence Tolerance = 1.0e-8
    """ % order
    ref = oracle.itersolve(A, b, method="bicgstab", precond="ilu%d" % order, ilu=F, tol=TOL, maxit=500)
    got = M.itersolver(b, np.zeros(A.n), sif)
    assert got["info"] == 1 and got["iters"] == ref["iters"]
    M.close()


@pytest.mark.parametrize("precond,restart", [("none", 10), ("diagonal", 5), ("ilu0", 5), ("ilu0", 30)])
def test_gmres_parity(oracle, b200, heat, heat_gpu, precond, restart):
    """GMRES(m), huti_dgmressolv (fhutiter/src/huti_gmres.F90:390-822), left-preconditioned as IterSolver calls it
    (IterSolve.F90:509-525): restart-cycle counts and solution against the oracle; nonsymmetric 4-dof system; SIF path."""
    A, b = heat
    ref = oracle.itersolve(A, b, method="gmres", precond=precond, tol=TOL, maxit=500, gmres_restart=restart)
    got = heat_gpu.solve(b, method="gmres", precond=precond, tol=TOL, maxit=500, gmres_restart=restart)
    assert got["info"] == ref["info"] == 1, (got["info"], ref["info"])
    assert iters_close(got["iters"], ref["iters"]), (got["iters"], ref["iters"])
    assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
    if precond == "ilu0" and restart == 5:
        A3, b3 = oracle.cavity_flow(6)
        A3 = A3.copy(); x = np.zeros(A3.n); oracle.scale_system(A3, b3, x)
        M = b200.Matrix(); M.set_structure(A3.rows, A3.cols, A3.diag, 1, 4); M.set_values(A3.vals)
        ref = oracle.itersolve(A3, b3, method="gmres", precond="ilu0", tol=TOL, maxit=300, gmres_restart=20)
        got = M.solve(b3, method="gmres", precond="ilu0", tol=TOL, maxit=300, gmres_restart=20)
        assert got["info"] == ref["info"] and iters_close(got["iters"], ref["iters"]), (got["info"], ref["info"], got["iters"], ref["iters"])
        if ref["info"] == 1:
            assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
        M.close()
        sif = """
          Linear System Solver = Iterative
          Linear System Iterative Method = GMRES
          Linear System GMRES Restart = 5
          Linear System Preconditioning = ILU0
          Linear System Max Iterations = 500
          Linear System Convergence Tolerance = 1.0e-8
        """
        ref = oracle.itersolve(A, b, method="gmres", precond="ilu0", tol=TOL, maxit=500, gmres_restart=5)
        got = heat_gpu.itersolver(b, None, sif, 0)
        assert got is not None and got["info"] == 1 and iters_close(got["iters"], ref["iters"])
        # maxiter: one cycle only
        ref = oracle.itersolve(A, b, method="gmres", precond="none", tol=1e-14, maxit=2, gmres_restart=3)
        got = heat_gpu.solve(b, method="gmres", precond="none", tol=1e-14, maxit=2, gmres_restart=3)
        assert got["info"] == ref["info"] == 2 and got["iters"] == ref["iters"]


def test_device_scaling(oracle, b200):
    """b200_scale_system = ScaleLinearSystemDiagonal on the device (SolverUtils.F90:12976-13213): scaled values bit-exact
    against the oracle's restatement; a solve on the UNSCALED b/x through the scaled handle returns the unscaled solution
    the reference obtains by scaling, solving and back-scaling (13515-13643), in the same number of iterations."""
    A0, b0 = oracle.heat_cube(16, faces=["x0", "y1"], source=3.0)
    rs = np.random.RandomState(31)
    sc = rs.uniform(0.5, 20.0, A0.n)                               # badly scaled, still symmetric: A <- S A S
    A0 = A0.copy(); A0.vals *= np.repeat(sc, np.diff(A0.rows)) * sc[A0.cols - 1]; b0 = b0 * sc
    A = A0.copy(); b = b0.copy(); x = np.zeros(A.n)
    D, bnorm = oracle.scale_system(A, b, x)
    M = b200.Matrix(); M.set_structure(A0.rows, A0.cols, A0.diag, 1, 1); M.set_values(A0.vals)
    M.scale_system()
    assert np.array_equal(M.values(), A.vals)
    for method, pc in [("bicgstab", "ilu0"), ("cg", "diagonal"), ("gmres", "ilu0")]:
        ref = oracle.itersolve(A, b, method=method, precond=pc, tol=TOL, maxit=500)
        xref = ref["x"] * D                                        # BackScale: x = x * Diag
        got = M.solve(b0, method=method, precond=pc, tol=TOL, maxit=500)
        assert got["info"] == ref["info"] == 1 and iters_close(got["iters"], ref["iters"]), (method, got["iters"], ref["iters"])
        assert rel_l2(got["x"], xref) <= 10 * TOL
    M.set_values(A0.vals)                                          # new values clear the scaled state
    assert np.array_equal(M.values(), A0.vals)
    M.close()


@pytest.mark.parametrize("method", ["cgs", "tfqmr", "bicgstab2"])
@pytest.mark.parametrize("precond", ["none", "diagonal", "ilu0"])
def test_cgs_tfqmr_parity(oracle, b200, heat, heat_gpu, method, precond):
    """huti_dcgssolv (fhutiter/src/huti_cgs.F90:283-470, right-oriented), huti_dtfqmrsolv (huti_tfqmr.F90:455-803) and
    huti_dbicgstab_2solv (huti_bicgstab_2.F90:339-578), the latter two left-oriented as IterSolver calls them: iteration counts and solutions against the oracle; nonsymmetric system; keyword path."""
    A, b = heat
    ref = oracle.itersolve(A, b, method=method, precond=precond, tol=TOL, maxit=500)
    got = heat_gpu.solve(b, method=method, precond=precond, tol=TOL, maxit=500)
    assert got["info"] == ref["info"] == 1, (got["info"], ref["info"])
    assert iters_close(got["iters"], ref["iters"]), (got["iters"], ref["iters"])
    assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
    if precond == "ilu0":
        A3, b3 = oracle.cavity_flow(6)
        A3 = A3.copy(); x = np.zeros(A3.n); oracle.scale_system(A3, b3, x)
        M = b200.Matrix(); M.set_structure(A3.rows, A3.cols, A3.diag, 1, 4); M.set_values(A3.vals)
        ref = oracle.itersolve(A3, b3, method=method, precond="ilu0", tol=TOL, maxit=300)
        got = M.solve(b3, method=method, precond="ilu0", tol=TOL, maxit=300)
        assert got["info"] == ref["info"], (got["info"], ref["info"])
        if ref["info"] == 1:
            assert iters_close(got["iters"], ref["iters"]) and rel_l2(got["x"], ref["x"]) <= 10 * TOL
        M.close()
        sif = """
          Linear System Solver = Iterative
          Linear System Iterative Method = %s
          Linear System Preconditioning = ILU0
          Linear System Max Iterations = 500
          Linear System Convergence Tolerance = 1.0e-8
        """ % {"cgs": "CGS", "tfqmr": "TFQMR", "bicgstab2": "BiCGStab2"}[method]
        ref = oracle.itersolve(A, b, method=method, precond="ilu0", tol=TOL, maxit=500)
        got = heat_gpu.itersolver(b, None, sif, 0)
        assert got is not None and got["info"] == 1 and iters_close(got["iters"], ref["iters"])


def test_reference_linearsolvers_case_gpu(oracle, b200):
    """fem/tests/linearsolvers/TempDist.sif through the C ABI: the reference's mesh, every Krylov method of the SIF + ILU0, tol 1e-12,
    device-side Linear System Scaling; answer = the constant k, `Reference Norm = k`; iteration counts as the oracle's."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from linearsolvers_case import tempdist_system, compute_norm
    from test_oracle_golden import LINSOLVERS
    for k, (method, kw) in enumerate(LINSOLVERS):
        k = float(k + 3)
        S, b, x0 = tempdist_system(k)
        A = oracle.CRS.from_scipy(S)
        M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 1); M.set_values(A.vals)
        M.scale_system()
        P = oracle.shadow_space(A.n, 4) if method == "idrs" else None
        got = M.solve(b, x0=x0, method=method, precond="ilu0", tol=1e-12, maxit=3500, P=P, **kw)
        ref = oracle.solve_linear_system(A, b, x0=x0, method=method, precond="ilu0", tol=1e-12, maxit=3500, P=P, **kw)
        assert got["info"] == ref["info"] == 1, (method, got["info"])
        assert abs(compute_norm(got["x"]) - k) <= 1e-5 * k and np.abs(got["x"] - k).max() <= 1e-9 * k, method
        assert iters_close(got["iters"], ref["iters"]), (method, got["iters"], ref["iters"])
        M.close()


def test_reference_winkel_poisson_norm_gpu(b200):
    """fem/tests/WinkelBmPoissonCgIlu0 / IdrsIlu0 through the C ABI (mesh by the reference's ElmerGrid, device-side scaling):
    `Reference Norm = 1.03281284`."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import winkel_case as W
    if not W.available():
        pytest.skip("oracle/_ref/ElmerGrid not built")
    A, b = W.system()
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 1); M.set_values(A.vals)
    M.scale_system()
    for method in ("cg", "idrs", "bicgstab", "gmres"):
        got = M.solve(b, method=method, precond="ilu0", tol=1e-8, maxit=1000)
        assert got["info"] == 1, method
        assert abs(W.norm(got["x"]) - W.REFERENCE_NORM) <= 1e-6 * W.REFERENCE_NORM, (method, W.norm(got["x"]))
    M.close()


def test_reference_winkel_navier_norm_gpu(oracle, b200):
    """fem/tests/WinkelBmNavier* through the C ABI: ndeg = 3 (block-column SpMV, 3 accumulators), device-side scaling, ILU0 / Jacobi:
    `Reference Norm = 2.25252433E-02`, SpMV and ILU0 bit-exact against the oracle on this unstructured-numbered operand."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import winkel_case as W
    if not W.available():
        pytest.skip("oracle/_ref/ElmerGrid not built")
    A, b = W.navier_system()
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 3); M.set_values(A.vals)
    u = np.random.RandomState(41).standard_normal(A.n)
    assert np.array_equal(M.matvec(u), oracle.matvec(A, u))
    M.scale_system()
    As = A.copy(); bs = b.copy(); xs = np.zeros(A.n); oracle.scale_system(As, bs, xs)
    M.factorize()
    assert np.array_equal(M.ilu_values(), oracle.ilu0(As))
    for method, pc in (("cg", "ilu0"), ("bicgstabl", "ilu0"), ("bicgstab", "diagonal"), ("gmres", "ilu0")):
        got = M.solve(b, method=method, precond=pc, tol=1e-10, maxit=5000, bicgstabl_l=4, gmres_restart=30)
        assert got["info"] == 1, method
        assert abs(W.norm(got["x"]) - W.NAVIER_REFERENCE_NORM) <= 1e-6 * W.NAVIER_REFERENCE_NORM, (method, W.norm(got["x"]))
    M.close()


def test_bilu_bit_exact(oracle, b200):
    """BILU ("Linear System Preconditioning = BILU", IterSolve.F90:549-558, 745-765): ILU0 of the block-diagonal part of the matrix
    (CRS_BlockDiagonal, CRSMatrix.F90:2382-2420), Blocks = dofs per node.  Pattern, values and triangular solves bit-exact against the
    oracle's ILU0 of the same block-diagonal matrix; Krylov iteration counts; keyword path."""
    import scipy.sparse as sp
    A, b = oracle.elasticity_beam(8, 3, 3, lx=2.5)
    A = A.copy(); x = np.zeros(A.n); oracle.scale_system(A, b, x)
    S = A.to_scipy().tocoo()
    keep = (S.row % 3) == (S.col % 3)
    Bm = oracle.CRS.from_scipy(sp.csr_matrix((S.data[keep], (S.row[keep], S.col[keep])), shape=S.shape))
    F = oracle.CRS(Bm.rows, Bm.cols, Bm.diag, oracle.ilu0(Bm), 1)
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 3); M.set_values(A.vals)
    M.set_bilu_blocks(3)
    r, c, d = M.ilu_structure()
    assert np.array_equal(r, F.rows) and np.array_equal(c, F.cols) and np.array_equal(d, F.diag)
    M.factorize()
    assert np.array_equal(M.ilu_values(), F.vals)
    v = np.random.RandomState(51).standard_normal(A.n)
    assert np.array_equal(M.lu_precondition(v), oracle.lu_precond(A, F, v))
    for method in ("cg", "bicgstabl"):
        ref = oracle.itersolve(A, b, method=method, precond="ilu0", ilu=F, tol=TOL, maxit=2000, bicgstabl_l=4)
        got = M.solve(b, method=method, precond="ilu0", tol=TOL, maxit=2000, bicgstabl_l=4)
        assert got["info"] == ref["info"] == 1 and iters_close(got["iters"], ref["iters"]), (method, got["iters"], ref["iters"])
        assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
    M.close()
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 3); M.set_values(A.vals)
    sif = """
      Linear System Solver = Iterative
      Linear System Iterative Method = CG
      Linear System Preconditioning = BILU
      Linear System Max Iterations = 2000
      Linear System Convergence Tolerance = 1.0e-8
    """
    ref = oracle.itersolve(A, b, method="cg", precond="ilu0", ilu=F, tol=TOL, maxit=2000)
    got = M.itersolver(b, None, sif, 0)
    assert got is not None and got["info"] == 1 and iters_close(got["iters"], ref["iters"])
    assert M.itersolver(b, None, sif.replace("BILU", "BILU1"), 0) is None          # order > 0: declined (see itersolver.cu)
    got0 = M.itersolver(b, None, sif.replace("BILU", "ILU0"), 0)                    # back to plain ILU0 on the same handle
    ref0 = oracle.itersolve(A, b, method="cg", precond="ilu0", tol=TOL, maxit=2000)
    assert got0["info"] == 1 and iters_close(got0["iters"], ref0["iters"])
    M.close()


def test_stationary_methods(oracle, b200):
    """itermethod_jacobi / itermethod_richardson (IterativeMethods.F90:297-521).  Jacobi on the reference's own linearsolvers case
    (TempDist.sif:71-76: `Reference Norm = 3`); Richardson (lumped-matrix scaling, meant for mass matrices) on a diagonally dominant system."""
    import sys, os
    import scipy.sparse as sp
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from linearsolvers_case import tempdist_system, compute_norm
    S, b, x0 = tempdist_system(3.0)
    A = oracle.CRS.from_scipy(S)
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 1); M.set_values(A.vals)
    M.scale_system()
    ref = oracle.solve_linear_system(A, b, x0=x0, method="jacobi", precond="none", tol=1e-12, maxit=3500)
    got = M.solve(b, x0=x0, method="jacobi", precond="none", tol=1e-12, maxit=3500)
    assert got["info"] == ref["info"] == 1 and iters_close(got["iters"], ref["iters"]), (got["iters"], ref["iters"])
    assert abs(compute_norm(got["x"]) - 3.0) <= 1e-5 * 3.0
    M.close()
    A2, _ = oracle.heat_cube(8, faces=["x0"], source=1.0)
    Sm = (sp.identity(A2.n) * 3.0 + 0.1 * abs(A2.to_scipy())).tocsr()       # mass-matrix-like: positive entries, dominant diagonal
    Am = oracle.CRS.from_scipy(Sm)
    bm = np.random.RandomState(61).standard_normal(Am.n)
    M = b200.Matrix(); M.set_structure(Am.rows, Am.cols, Am.diag, 1, 1); M.set_values(Am.vals)
    for method in ("richardson", "jacobi"):
        ref = oracle.itersolve(Am, bm, method=method, precond="none", tol=1e-10, maxit=500)
        got = M.solve(bm, method=method, precond="none", tol=1e-10, maxit=500)
        assert got["info"] == ref["info"] == 1 and got["iters"] == ref["iters"], (method, got["iters"], ref["iters"])
        assert rel_l2(got["x"], ref["x"]) <= 1e-9
    sif = """
      Linear System Solver = Iterative
      Linear System Iterative Method = Richardson
      Linear System Max Iterations = 500
      Linear System Convergence Tolerance = 1.0e-10
    """
    got = M.itersolver(bm, None, sif, 0)
    assert got is not None and got["info"] == 1
    M.close()


def test_sgs(oracle, b200, heat, heat_gpu):
    """itermethod_sgs (IterativeMethods.F90:179-285): the Gauss-Seidel wavefront kernel reproduces the reference's sequential sweeps bit for
    bit (x after 1 and 3 rounds); converged solves: round counts as the oracle; the reference's linearsolvers case (TempDist.sif:80-85,
    `Reference Norm = 4`); keyword path with the default `SGS Overrelaxation Factor` (the REAL literal 1.8, IterSolve.F90:358)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from linearsolvers_case import tempdist_system, compute_norm
    A, b = heat
    for rounds in (1, 3):
        ref = oracle.itersolve(A, b, method="sgs", precond="none", tol=1e-30, maxit=rounds)
        got = heat_gpu.solve(b, method="sgs", precond="none", tol=1e-30, maxit=rounds)
        assert got["info"] == ref["info"] == 2 and got["iters"] == ref["iters"] == rounds
        assert np.array_equal(got["x"], ref["x"]), rounds
    ref = oracle.itersolve(A, b, method="sgs", precond="none", tol=TOL, maxit=2000, sgs_omega=1.5)
    got = heat_gpu.solve(b, method="sgs", precond="none", tol=TOL, maxit=2000, sgs_omega=1.5)
    assert got["info"] == ref["info"] == 1 and got["iters"] == ref["iters"]
    assert rel_l2(got["x"], ref["x"]) <= 10 * TOL
    A3, b3 = oracle.cavity_flow(4)                          # nonsymmetric, 4 dofs per node: sweeps still bit-exact
    A3 = A3.copy(); x = np.zeros(A3.n); oracle.scale_system(A3, b3, x)
    M = b200.Matrix(); M.set_structure(A3.rows, A3.cols, A3.diag, 1, 4); M.set_values(A3.vals)
    ref = oracle.itersolve(A3, b3, method="sgs", precond="none", tol=1e-30, maxit=2, sgs_omega=1.0)
    got = M.solve(b3, method="sgs", precond="none", tol=1e-30, maxit=2, sgs_omega=1.0)
    assert np.array_equal(got["x"], ref["x"])
    M.close()
    S, bb, x0 = tempdist_system(4.0)
    At = oracle.CRS.from_scipy(S)
    M = b200.Matrix(); M.set_structure(At.rows, At.cols, At.diag, 1, 1); M.set_values(At.vals)
    M.scale_system()
    sif = """
      Linear System Solver = Iterative
      Linear System Iterative Method = SGS
      Linear System Max Iterations = 3500
      Linear System Convergence Tolerance = 1.0e-12
    """
    ref = oracle.solve_linear_system(At, bb, x0=x0, method="sgs", precond="none", tol=1e-12, maxit=3500)
    got = M.itersolver(bb, x0, sif, 0)
    assert got is not None and got["info"] == ref["info"] == 1 and iters_close(got["iters"], ref["iters"]), (got["iters"], ref["iters"])
    assert abs(compute_norm(got["x"]) - 4.0) <= 1e-5 * 4.0
    M.close()


@pytest.mark.parametrize("ndeg", [2, 5, 6, 8, 10])
def test_spmv_ndeg_variants_bit_exact(oracle, b200, ndeg):
    """CRS_MatrixVectorProd's ndeg loops that no benchmark operand exercises (CRSMatrix.F90:4794-4856): ndeg = 2 -> 2 partial sums,
    5 and 10 -> 5, 6 -> 3, 8 -> 4; one column index per group of consecutive columns.  Node-block matrices (ndeg interleaved dofs per node
    of a hex8 grid, dense ndeg x ndeg blocks, random values): the device product equals the oracle's bit for bit, and scipy's to rounding."""
    from elmerfem_b200 import synth
    xyz, elems = synth.grid_hex8(5, 4, 3)
    rows, cols, diag = synth.crs_structure(xyz.shape[0], elems, ndeg)
    rs = np.random.RandomState(100 + ndeg)
    A = synth.CRS(rows, cols, diag, rs.standard_normal(cols.size), ndeg)
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, ndeg); M.set_values(A.vals)
    for _ in range(2):
        u = rs.standard_normal(A.n)
        got, ref = M.matvec(u), oracle.matvec(A, u)
        assert np.array_equal(got, ref), ndeg
        sref = A.to_scipy() @ u
        assert np.abs(got - sref).max() <= 1e-12 * np.abs(sref).max()
    M.close()


@pytest.mark.parametrize("ndeg", [2, 3, 4, 5])
@pytest.mark.parametrize("node_u", ["0", "1"])
def test_node_lane_sweeps_bit_exact(oracle, b200, monkeypatch, ndeg, node_u):
    """Node-lane triangular solves (structure.cu node_lane_layout, precond.cu k_sptrsv_wide_node) forced on (B200_TRI_NODE=2) for every
    dof count they are instantiated for: node-block matrices with dense ndeg x ndeg blocks, ILU0 factor and both sweeps bit-identical to
    CRS_IncompleteLU / CRS_LUSolve (CRSMatrix.F90:3604-3660, 4642-4660), with the backward plan in the node-lane and in the row-level layout."""
    from elmerfem_b200 import synth
    monkeypatch.setenv("B200_TRI_NODE", "2")
    monkeypatch.setenv("B200_TRI_NODE_U", node_u)
    ex = 2 if ndeg >= 5 else (3 if ndeg == 4 else 4)              # rows of at most 64 entries per triangle: 27 * ndeg / 2
    xyz, elems = synth.grid_hex8(ex + 5, ex + 1, ex)
    if 27 * ndeg > 128:                                           # 5 and 6 dofs: a bar of single elements (12-node stencils) keeps the rows narrow enough
        xyz, elems = synth.grid_hex8(14, 1, 1)
    rows, cols, diag = synth.crs_structure(xyz.shape[0], elems, ndeg)
    rs = np.random.RandomState(200 + ndeg)
    vals = 0.05 * rs.standard_normal(cols.size)
    vals[diag - 1] = 1.0 + rs.random_sample(diag.size)
    A = synth.CRS(rows, cols, diag, vals, ndeg)
    assert int(np.max(np.maximum(A.diag - A.rows[:-1], A.rows[1:] - A.diag - 1))) <= 64
    M = b200.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, ndeg); M.set_values(A.vals)
    try:
        M.factorize()
        ilu = oracle.ilu0(A)
        assert np.array_equal(M.ilu_values(), ilu)
        lv = M.levels()
        nn = A.n // ndeg
        assert lv["forward"] < A.n and lv["forward"] <= nn       # node levels, not row levels
        for seed in (1, 2):
            v = rs.standard_normal(A.n)
            if seed == 2:
                v[::4] = 0.0
            got, ref = M.lu_precondition(v), oracle.lu_precond(A, ilu, v)
            assert np.array_equal(got.view(np.int64), ref.view(np.int64)), (ndeg, node_u)
        b = rs.standard_normal(A.n)
        ref = oracle.itersolve(A, b, method="bicgstab", precond="ilu0", tol=1e-10, maxit=200)
        got = M.solve(b, method="bicgstab", precond="ilu0", tol=1e-10, maxit=200)
        assert got["info"] == ref["info"] and abs(got["iters"] - ref["iters"]) <= 1
    finally:
        M.close()

"""`Linear System Robust` (IterSolve.F90:482-496; BiCGStab(l): IterativeMethods.F90:1110-1139, IDR(s): 1862-1898) on the
oracle, CPU only.  The reference has no golden vector for this keyword ("parity unpinned" for it); what is checked
here is the restated logic against its own definition: the best iterate is what comes back, it is declared
converged when it beats the robust tolerance, and nothing changes when the safeguard never fires."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def cavity(oracle):
    A, b = oracle.cavity_flow(6)
    A = A.copy()
    oracle.scale_system(A, b, np.zeros(A.n))
    return A, b


def _true_res(oracle, A, b, x):
    return float(np.linalg.norm(oracle.matvec(A, x) - b) / np.linalg.norm(b))


CASES = [("bicgstabl", dict(bicgstabl_l=2), "none"), ("bicgstabl", dict(bicgstabl_l=2), "ilu0"),
         ("bicgstabl", dict(bicgstabl_l=4), "none"), ("idrs", dict(idrs_s=4), "none")]


@pytest.mark.parametrize("method,kw,pc", CASES)
def test_robust_returns_the_best_iterate(oracle, cavity, method, kw, pc):
    A, b = cavity
    base = oracle.itersolve(A, b, method=method, precond=pc, tol=1e-10, maxit=400, **kw)
    rb = oracle.itersolve(A, b, method=method, precond=pc, tol=1e-10, maxit=400, robust=True, robust_tol=1e-4,
                          robust_limit=1e-2, robust_max_bad=0, **kw)
    assert base["info"] == 1 and rb["info"] == 1            # BestNorm < Robust Tolerance => Converged
    assert rb["iters"] < base["iters"]                      # stopped at the first non-improving step below 1e-4
    true = _true_res(oracle, A, b, rb["x"])
    assert true < 1e-4                                      # what came back beats the robust tolerance ...
    assert true <= rb["residual"] * (1 + 1e-6)              # ... and is not worse than the iterate it stopped at
    assert rb["residual"] > 1e-10                           # and the plain tolerance was NOT reached


@pytest.mark.parametrize("method,kw,pc", CASES)
def test_robust_with_room_is_the_plain_solve(oracle, cavity, method, kw, pc):
    """`Robust Max Iterations` generous and `Robust Limit` never exceeded: same iterations, same bits."""
    A, b = cavity
    base = oracle.itersolve(A, b, method=method, precond=pc, tol=1e-10, maxit=400, **kw)
    rb = oracle.itersolve(A, b, method=method, precond=pc, tol=1e-10, maxit=400, robust=True, robust_tol=1e-4,
                          robust_limit=1e20, robust_max_bad=400, **kw)
    assert rb["iters"] == base["iters"] and rb["info"] == base["info"]
    assert np.array_equal(rb["x"], base["x"])


def test_robust_defaults_follow_itersolve(oracle):
    ipar, dpar = oracle.fill_ipar_dpar(10, "idrs", tol=1e-9, maxit=77, robust=True)
    assert ipar[26 - 1] == 1 and ipar[27 - 1] == 38 and ipar[29 - 1] == 1        # MAXIT/2, start 1
    assert dpar[4 - 1] == 1.1 and dpar[5 - 1] == np.sqrt(1e-9)
    assert dpar[3 - 1] == 1e-9 ** float(np.float32(2.0) / np.float32(3.0))       # HUTI_TOLERANCE**(2.0/3.0), default-real exponent
    import elmerfem_b200 as B
    ip2, dp2 = B.fill_ipar_dpar(10, "idrs", tol=1e-9, maxit=77, robust=True)
    assert np.array_equal(ipar, ip2) and np.array_equal(dpar, dp2)


def test_robust_gives_up_as_maxiter_when_nothing_good_was_seen(oracle, cavity):
    A, b = cavity
    rb = oracle.itersolve(A, b, method="bicgstabl", precond="none", bicgstabl_l=2, tol=1e-10, maxit=5, robust=True,
                          robust_tol=1e-6, robust_limit=1e-2, robust_max_bad=0)
    assert rb["info"] == 2 and rb["iters"] == 5             # HUTI_MAXITER: BestNorm never got below the robust tolerance

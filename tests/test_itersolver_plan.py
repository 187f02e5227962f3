"""IterSolver's keyword decisions (IterSolve.F90:250-577) as the library takes them -- b200_itersolver_plan, host-only -- against
the oracle's restatement of the same ipar/dpar filling, on keyword sections written the way the reference's test SIFs write them
(type prefixes, quotes, comments, mixed case, `Solver 1 ::` prefixes).  CPU only."""
import numpy as np
import pytest

import elmerfem_b200 as B

METHOD = {"cg": 1, "bicgstab": 2, "bicgstabl": 3, "gcr": 4, "idrs": 5, "gmres": 6, "cgs": 7, "tfqmr": 8, "bicgstab2": 9,
          "jacobi": 10, "richardson": 11, "sgs": 12}


def _plan(sif, n=1000, ndeg=1):
    return B.itersolver_plan(sif, n, ndeg)


@pytest.mark.parametrize("method", sorted(METHOD))
def test_defaults_match_the_oracle_filling(oracle, method):
    sif = """
      Linear System Solver = Iterative
      Linear System Iterative Method = %s
      Linear System Max Iterations = 350
      Linear System Convergence Tolerance = 1.0e-9
      Linear System Preconditioning = ILU0
    """ % method
    p = _plan(sif)
    assert p is not None and p["method"] == METHOD[method] and p["precond"] == 2 and p["ilu_order"] == 0
    ipar, dpar = oracle.fill_ipar_dpar(1000, method, tol=1e-9, maxit=350, residual_output=1)
    assert np.array_equal(p["ipar"], ipar), np.nonzero(p["ipar"] != ipar)
    ok = np.ones(10, dtype=bool)
    if method != "sgs":
        ok[2] = False                      # dpar(3) is the SGS factor slot; the oracle helper fills it for every method
    assert np.array_equal(p["dpar"][ok], dpar[ok])


def test_method_parameters_and_robust(oracle):
    sif = """
      linear system iterative method = "BiCGStabL"     ! quoted, mixed case
      BiCGstabl Polynomial Degree = Integer 4
      Linear System Max Iterations = Integer 77
      Linear System Convergence Tolerance = Real 1.0d-9
      Linear System Residual Output = 0
      Linear System Preconditioning = Diagonal
      Linear System Robust = Logical True
    """
    p = _plan(sif)
    ipar, dpar = oracle.fill_ipar_dpar(1000, "bicgstabl", tol=1e-9, maxit=77, residual_output=0, bicgstabl_l=4, robust=True)
    assert p["method"] == 3 and p["precond"] == 1 and p["ilu_order"] == -1
    assert np.array_equal(p["ipar"], ipar) and np.array_equal(p["dpar"], dpar)
    # explicit robust keywords
    p = _plan(sif + """
      Linear System Robust Tolerance = 1e-4
      Linear System Robust Limit = 1e-2
      Linear System Robust Margin = 1.25
      Linear System Robust Max Iterations = 3
      Linear System Robust Start Iteration = 5
    """)
    ipar, dpar = oracle.fill_ipar_dpar(1000, "bicgstabl", tol=1e-9, maxit=77, residual_output=0, bicgstabl_l=4, robust=True,
                                       robust_tol=1e-4, robust_limit=1e-2, robust_margin=1.25, robust_max_bad=3, robust_start=5)
    assert np.array_equal(p["ipar"], ipar) and np.array_equal(p["dpar"], dpar)
    for method, extra, kw in [("gcr", "Linear System GCR Restart = 30", dict(gcr_restart=30)), ("gcr", "", dict()),
                              ("idrs", "IDRS Parameter = 6\nIDRS Smoothing = True", dict(idrs_s=6, smoothing=True)),
                              ("gmres", "Linear System GMRES Restart = 25", dict(gmres_restart=25)),
                              ("sgs", "SGS Overrelaxation Factor = 1.5", dict(sgs_omega=1.5))]:
        p = _plan("Linear System Iterative Method = %s\nLinear System Max Iterations = 500\n"
                  "Linear System Convergence Tolerance = 1e-8\nLinear System Min Iterations = 2\n%s\n" % (method, extra))
        ipar, dpar = oracle.fill_ipar_dpar(1000, method, tol=1e-8, maxit=500, minit=2, residual_output=1, **kw)
        assert np.array_equal(p["ipar"], ipar), method
        if method == "sgs":
            assert p["dpar"][2] == 1.5


@pytest.mark.parametrize("text,order,blocks", [("ILU0", 0, 0), ("ilu", 0, 0), ("ILU1", 1, 0), ("Ilu3", 3, 0), ("ILU9", 9, 0),
                                               ("BILU", 0, 3), ("bilu0", 0, 3), ("None", -1, 0), ("Diagonal", -1, 0)])
def test_preconditioner_selection(text, order, blocks):
    p = _plan("Linear System Max Iterations = 10\nLinear System Preconditioning = %s\n" % text, ndeg=3)
    assert p["ilu_order"] == order and p["bilu_blocks"] == blocks
    assert p["precond"] == (2 if order >= 0 else (1 if text.lower() == "diagonal" else 0))
    assert p["method"] == 2                                   # absent method keyword: BiCGStab (IterSolve.F90:251-253)


def test_ilu_order_keyword_wins():
    p = _plan("Linear System Max Iterations = 10\nLinear System Preconditioning = ILU0\nLinear System ILU Order = 2\n")
    assert p["ilu_order"] == 2


def test_the_reference_sif_sections():
    # fem/tests/CoordinateScaling/case.sif, Solver 1
    p = _plan("""
      Linear System Solver = Iterative
      Linear System Iterative Method = BiCGStab
      Linear System Max Iterations = 500
      Linear System Convergence Tolerance = 1.0e-8
      Linear System Preconditioning = ILU1
      Linear System ILUT Tolerance = 1.0e-3
      Linear System Abort Not Converged = False
      Linear System Residual Output = 10
      Linear System Precondition Recompute = 1
    """, n=441)
    assert (p["method"], p["precond"], p["ilu_order"]) == (2, 2, 1)
    assert p["ipar"][2] == 441 and p["ipar"][9] == 500 and p["ipar"][4] == 10 and p["dpar"][0] == 1e-8 and p["dpar"][1] == 1e20
    # fem/tests/WinkelBmPoissonIdrsIlu0/case.sif style, with the `Solver 1 ::` prefix form
    p = _plan("Solver 1 :: Linear System Iterative Method = Idrs\nSolver 1 :: Linear System Max Iterations = 1000\n"
              "Solver 1 :: Idrs Parameter = 4\nSolver 1 :: Linear System Preconditioning = ILU0\n")
    assert (p["method"], p["precond"], p["ipar"][17]) == (5, 2, 4)


@pytest.mark.parametrize("line,why", [
    ("Linear System Complex = True", "complex"),
    ("Linear System Preconditioning = BILU2", "BILU order"),
    ("Linear System Preconditioning = Multigrid", "multigrid"),
    ("Linear System Preconditioning = vanka", "vanka"),
    ("Linear System Normwise Backward Error = True", "backward-error"),
    ("Linear System Componentwise Backward Error = True", "backward-error"),
    ("Linear System ILU Factor = 0.1", "ILU Factor"),
    ("Edge Basis = True", "Edge Basis"),
    ("Linear System Iterative Method = CG\nLinear System Left Preconditioning = True", "left"),
])
def test_declined_combinations(line, why):
    assert _plan("Linear System Max Iterations = 10\n" + line + "\n") is None
    assert why.lower() in B.last_error().lower()


def test_symmetric_ilu_is_accepted():
    """'Linear System Symmetric ILU' (A % Cholesky, IterSolve.F90:526) no longer declines: the incomplete Cholesky branches are built."""
    p = _plan("Linear System Max Iterations = 10\nLinear System Iterative Method = CG\nLinear System Preconditioning = ILU0\n"
              "Linear System Symmetric ILU = True\n")
    assert p is not None and (p["method"], p["precond"], p["ilu_order"]) == (1, 2, 0)


def test_left_preconditioning_is_kept_where_itersolver_does_it():
    for method in ("gmres", "tfqmr", "bicgstab2", "bicgstabl"):
        assert _plan("Linear System Max Iterations = 10\nLinear System Iterative Method = %s\n"
                     "Linear System Left Preconditioning = True\n" % method) is not None


def test_errors_are_errors_not_declines():
    with pytest.raises(B.B200Error):
        _plan("Linear System Iterative Method = CG\n")                                  # Max Iterations missing
    with pytest.raises(B.B200Error):
        _plan("Linear System Max Iterations = 10\nLinear System Iterative Method = bicgstabl\nBiCGstabl polynomial degree = 1\n")


def test_unknown_method_name_runs_bicgstab_as_in_itersolve():
    p = _plan("Linear System Max Iterations = 10\nLinear System Iterative Method = FancyNewMethod\n")   # CASE DEFAULT, IterSolve.F90:313-314
    assert p["method"] == 2 and p["ipar"][3] == 8

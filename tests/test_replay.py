"""The replay tool (elmerfem_b200/replay.py): `Linear System Save` dump + the user's .sif -> IterSolver's plan (host-only) and, on a
GPU box, the solve through the C ABI.  CPU test: SIF section extraction, plan, decline, and that a solve without a GPU fails loudly
instead of falling back to anything."""
import ctypes as C
import os

import numpy as np
import pytest

import elmerfem_b200 as B
from elmerfem_b200 import meshio, replay, synth

SIF = """
Header
  Mesh DB "." "cube"
End
Solver 2
  Equation = SaveScalars      ! not a linear solve
  Procedure = "SaveData" "SaveScalars"
End
Solver 1
  Equation = HeatSolver
  Variable = Temperature
  Linear System Solver = Iterative
  Linear System Iterative Method = BiCGStabl   ! comment after the value
  BiCGstabl polynomial degree = 4
  Linear System Max Iterations = 500
  Linear System Convergence Tolerance = 1.0e-9
  Linear System Preconditioning = ILU1
  Linear System Save = True
End
Solver 1 :: Reference Norm = 0.123
Solver 1 :: Linear System Residual Output = 0
Solver 3
  Linear System Solver = Iterative
  Linear System Iterative Method = GCR
  Linear System Max Iterations = 10
  Linear System Preconditioning = Multigrid
End
"""


@pytest.fixture()
def dump(tmp_path):
    A, b = synth.heat_cube(4, faces=["x0"])
    meshio.write_linsys(A.to_scipy(), b, "linsys", str(tmp_path))
    sif = tmp_path / "case.sif"
    sif.write_text(SIF)
    return str(tmp_path), str(sif), A.n


def test_solver_section_extraction():
    sec = replay.solver_section(SIF, 1)
    assert "BiCGStabl" in sec and "SaveScalars" not in sec and "GCR" not in sec
    assert "Linear System Residual Output = 0" in sec and "comment" not in sec
    plan = B.itersolver_plan(sec, 125)
    assert (plan["method"], plan["precond"], plan["ilu_order"]) == (3, 2, 1)
    assert plan["ipar"][15] == 4 and plan["ipar"][4] == 0 and plan["dpar"][0] == 1e-9


def test_plan_only_and_decline(dump, capsys):
    d, sif, n = dump
    assert replay.main(["--dir", d, "--sif", sif, "--plan-only"]) == 0
    out = capsys.readouterr().out
    assert "n = %d" % n in out and "method bicgstabl" in out and "ilu1" in out and "tolerance 1.000e-09" in out
    assert replay.main(["--dir", d, "--sif", sif, "--solver", "3", "--plan-only"]) == 3
    assert "DECLINED" in capsys.readouterr().out


def test_solve_needs_a_gpu(dump, capsys):
    d, sif, n = dump
    cnt = C.c_int(0)
    have_gpu = B.lib().b200_device_count(C.byref(cnt)) == 0 and cnt.value > 0
    if have_gpu:
        assert replay.main(["--dir", d, "--sif", sif]) == 0
        assert "HUTI_INFO = 1" in capsys.readouterr().out
    else:
        with pytest.raises(B.B200Error):
            replay.main(["--dir", d, "--sif", sif])


@pytest.mark.gpu
def test_replay_on_the_gpu_matches_the_oracle(dump, capsys, oracle):
    """The whole tool on hardware: dump -> reader -> keyword front-end -> device scaling -> BiCGStab(4)+ILU(1) -> back-scaling.  Iteration
    count as the oracle under the device summation order; solution norm and residual as the oracle's (SolveLinearSystem's order of
    operations: ScaleLinearSystem, IterSolver, BackScaleLinearSystem)."""
    import re
    d, sif, n = dump
    assert replay.main(["--dir", d, "--sif", sif]) == 0
    out = capsys.readouterr().out
    m = re.search(r"HUTI_INFO = (\d+), iterations = (\d+), \|\|Ax-b\|\|/\|\|b\|\| = ([0-9.eE+-]+), norm = ([0-9.eE+-]+)", out)
    assert m and int(m.group(1)) == 1
    S, b = meshio.read_linsys("linsys", d)
    A = synth.CRS.from_scipy(S, 1)
    oracle.set_dot_order(3)
    try:
        ref = oracle.solve_linear_system(A, b, method="bicgstabl", precond="ilu1", tol=1e-9, maxit=500, bicgstabl_l=4)
    finally:
        oracle.set_dot_order(0)
    assert ref["info"] == 1 and int(m.group(2)) == ref["iters"]
    assert float(m.group(3)) < 1e-8
    x = ref["x"]
    assert abs(float(m.group(4)) - float(np.sqrt(np.sum(x * x) / x.size))) <= 1e-7 * float(m.group(4))

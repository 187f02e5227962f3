"""Replay of a linear system saved by Elmer (`Linear System Save = True`: linsys_a.dat / linsys_b.dat, written by
SaveLinearSystem, fem/src/SolverUtils.F90:20313-20446) through the C ABI, with the Solver section of the user's own .sif:

    python -m elmerfem_b200.replay --dir RUN_DIR --sif case.sif [--solver 1] [--prefix linsys] [--no-scaling] [--plan-only]

prints what IterSolver decides from the keywords (b200_itersolver_plan, host-only) and, on a machine with a B200, solves
the system the way SolveLinearSystem would (default diagonal scaling on the device, IterSolver, back-scaling) and reports
HUTI_INFO, the iteration count, the true residual and ComputeNorm's norm -- the numbers to hold against the log of the
Elmer run that wrote the dump.  There is no CPU solve: without a GPU only --plan-only works."""
import argparse
import re
import sys

import numpy as np

import elmerfem_b200 as B
from elmerfem_b200 import meshio


def solver_section(sif_text, number=1):
    """Text of `Solver <number> ... End` plus every `Solver <number> :: key = value` line elsewhere in the file."""
    lines = sif_text.splitlines()
    out, inside = [], False
    head = re.compile(r"^\s*solver\s+(\d+)\s*$", re.I)
    for raw in lines:
        line = raw.split("!")[0].rstrip()
        m = head.match(line)
        if m:
            inside = int(m.group(1)) == number
            continue
        if inside and re.match(r"^\s*end\s*$", line, re.I):
            inside = False
            continue
        if inside:
            out.append(line)
        else:
            m2 = re.match(r"^\s*solver\s+(\d+)\s*::\s*(.*)$", line, re.I)
            if m2 and int(m2.group(1)) == number:
                out.append(m2.group(2))
    return "\n".join(l for l in out if "=" in l) + "\n"


def describe_plan(plan):
    names = {v: k for k, v in B.METHODS.items()}
    pcs = {0: "none", 1: "diagonal", 2: "ilu"}
    pc = pcs[plan["precond"]]
    if plan["precond"] == 2:
        pc = ("bilu0 (%d blocks)" % plan["bilu_blocks"]) if plan["bilu_blocks"] > 1 else "ilu%d" % plan["ilu_order"]
    return "method %s, preconditioner %s, max iterations %d, tolerance %.3e" % (
        names.get(plan["method"], "?"), pc, int(plan["ipar"][9]), float(plan["dpar"][0]))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--dir", default=".")
    ap.add_argument("--prefix", default="linsys")
    ap.add_argument("--sif", required=True)
    ap.add_argument("--solver", type=int, default=1)
    ap.add_argument("--ndeg", type=int, default=1, help="dofs per node (Matrix_t % ndeg / Solver % Variable % Dofs)")
    ap.add_argument("--no-scaling", action="store_true", help="Linear System Scaling = False")
    ap.add_argument("--plan-only", action="store_true")
    a = ap.parse_args(argv)
    section = solver_section(open(a.sif).read(), a.solver)
    S, b = meshio.read_linsys(a.prefix, a.dir)
    n = S.shape[0]
    plan = B.itersolver_plan(section, n, a.ndeg)
    if plan is None:
        print("DECLINED by the accelerated path (%s): Elmer's own solver would run" % B.last_error())
        return 3
    print("n = %d, nnz = %d: %s" % (n, S.nnz, describe_plan(plan)))
    if a.plan_only:
        return 0
    from elmerfem_b200 import synth
    A = synth.CRS.from_scipy(S, a.ndeg)
    M = B.Matrix()                                      # raises B200Error without a CUDA device: no CPU fallback
    try:
        M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg)
        M.set_values(A.vals)
        if not a.no_scaling:
            M.scale_system()
        out = M.itersolver(b, None, section)
        x = out["x"]
        res = float(np.linalg.norm(S @ x - b) / max(np.linalg.norm(b), 1e-300))
        print("HUTI_INFO = %d, iterations = %d, ||Ax-b||/||b|| = %.3e, norm = %.8E" %
              (out["info"], out["iters"], res, float(np.sqrt(np.sum(x * x) / x.size))))
        return 0 if out["info"] == 1 else 1
    finally:
        M.close()


if __name__ == "__main__":
    sys.exit(main())

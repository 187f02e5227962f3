"""Short driver for ncu captures: heat 200^3 (BASELINE configs[1]) -> a few SpMV, ILU0 applications and
one short BiCGStab+ILU0 solve.  Run under `ncu` on one GPU (see profiles/README.md for the commands)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import elmerfem_b200 as B          # noqa: E402
from elmerfem_b200 import synth    # noqa: E402

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 200
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 5
A, b = synth.workload("heat", ne)
M = B.Matrix()
M.set_structure(A.rows, A.cols, A.diag, 1, 1)
M.set_values(A.vals)
M.factorize()
print("spmv ms", M.time_matvec(3), "lu ms", M.time_lu(2))
r = M.solve(b, method="bicgstab", precond="ilu0", tol=1e-8, maxit=maxit)
print("solve", r["info"], r["iters"], r["stats"]["solve_ms"])
M.close()

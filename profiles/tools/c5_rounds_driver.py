"""ncu driver: a few BiCGStab(4)+Jacobi rounds on the C5 elasticity operand (one GPU's share): launch list of one round."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import elmerfem_b200 as B
from elmerfem_b200 import synth
nz = int(sys.argv[1]) if len(sys.argv) > 1 else 139
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
p = synth.elasticity_slab(137, 137, nz, 0, 1)
n = p["rows"].size - 1
rowid = np.repeat(np.arange(1, n + 1, dtype=np.int64), np.diff(p["rows"]))
diag = (np.flatnonzero(p["cols"] == rowid) + 1).astype(np.int32)
M = B.Matrix(); M.set_structure(p["rows"], p["cols"], diag, 1, 3); M.set_values(p["vals"])
r = M.solve(p["b"], method="bicgstabl", precond="diagonal", tol=1e-30, maxit=rounds, bicgstabl_l=4)
print("n", n, "rounds", r["iters"], "solve_ms", r["stats"]["solve_ms"], "ms/round", r["stats"]["solve_ms"] / max(r["iters"], 1))
M.close()

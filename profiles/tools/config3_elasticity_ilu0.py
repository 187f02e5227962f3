"""BASELINE configs[2] on one GPU's share: 3-dof elasticity, ~8 M dofs, BiCGStab(l=4) + ILU0: factor, LU application, rounds/s."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import elmerfem_b200 as B
from elmerfem_b200 import synth
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 137
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 140
p = synth.elasticity_slab(ne, ne, nz - 1, 0, 1)
n = p["rows"].size - 1
rowid = np.repeat(np.arange(1, n + 1, dtype=np.int64), np.diff(p["rows"]))
diag = (np.flatnonzero(p["cols"] == rowid) + 1).astype(np.int32)
M = B.Matrix(); M.set_structure(p["rows"], p["cols"], diag, 1, 3); M.set_values(p["vals"])
t = time.time(); M.factorize(); tf = time.time() - t
lv = M.levels()
print("elasticity %dx%dx%d: n %d nnz %d, levels %d/%d, first factorize %.2f s (factor %.1f ms), LU apply %.3f ms, SpMV %.3f ms" %
      (ne, ne, nz - 1, n, p["cols"].size, lv["forward"], lv["backward"], tf, M.stats()["factor_ms"], M.time_lu(5), M.time_matvec(10)), flush=True)
g = M.solve(p["b"], method="bicgstabl", precond="ilu0", tol=1e-8, maxit=int(sys.argv[3]) if len(sys.argv) > 3 else 20, bicgstabl_l=4)
print("BiCGStab(4)+ILU0: rounds %d info %d solve %.1f ms -> %.2f rounds/s" % (g["iters"], g["info"], g["stats"]["solve_ms"], g["iters"] / g["stats"]["solve_ms"] * 1e3), flush=True)
M.close()

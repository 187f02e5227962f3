"""BASELINE configs[0]: heat 100^3 (1.03 M dofs), CG + Jacobi: iterations/s on the device (resident solve)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import elmerfem_b200 as B
from elmerfem_b200 import synth
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
A, b = synth.workload("heat", ne)
M = B.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 1); M.set_values(A.vals)
for method, pc in (("cg", "diagonal"), ("cg", "ilu0"), ("bicgstab", "ilu0"), ("gmres", "ilu0"), ("gcr", "ilu0"), ("idrs", "ilu0"), ("bicgstabl", "ilu0")):
    for rep in range(2):
        g = M.solve(b, method=method, precond=pc, tol=1e-8, maxit=3000, bicgstabl_l=4, gmres_restart=30)
    st = g["stats"]
    print("heat %d^3 %-10s %-8s iters %4d info %d solve %.2f ms -> %.0f it/s, %d launches (%.1f us per launch), spmv %.3f ms" %
          (ne, method, pc, g["iters"], g["info"], st["solve_ms"], g["iters"] / st["solve_ms"] * 1e3, st["launches"], st["solve_ms"] * 1e3 / max(1, st["launches"]), M.time_matvec(20)), flush=True)
M.close()

import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import elmerfem_b200 as B
from elmerfem_b200 import synth
A, b = synth.workload("heat", 200)
for order in (0, 1):
    M = B.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 1); M.set_values(A.vals)
    M.set_ilu_order(order)
    t = time.time(); M.factorize(); tf = time.time() - t
    r, c, d = M.ilu_structure()
    lu = M.time_lu(5)
    g = M.solve(b, method="bicgstab", precond="ilu0", tol=1e-8, maxit=2000)
    st = g["stats"]
    print("heat 200^3 ILU(%d): pattern %d entries, first factorize (incl. symbolic + plans) %.2f s, factor %.1f ms, LU apply %.3f ms, levels %s, BiCGStab iters %d info %d solve %.1f ms" %
          (order, c.size, tf, st["factor_ms"], lu, M.levels(), g["iters"], g["info"], st["solve_ms"]), flush=True)
    M.close()

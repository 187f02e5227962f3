import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
from oracle import oracle as O
from elmerfem_b200 import synth
A, b = synth.workload("heat", 200)
Ao = O.CRS(A.rows, A.cols, A.diag, A.vals, 1)
O.set_threads(O.max_threads())
t = time.time(); r = O.itersolve(Ao, b, method="bicgstab", precond="ilu0", tol=1e-8, maxit=2000)
print("oracle C2 full size: info", r["info"], "iters", r["iters"], "residual", r.get("residual"), "time %.1f s" % (time.time() - t), flush=True)
np.save("gpurun_out/oracle_c2_x.npy", r["x"][::997])
import elmerfem_b200 as B
M = B.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 1); M.set_values(A.vals)
g = M.solve(b, method="bicgstab", precond="ilu0", tol=1e-8, maxit=2000)
print("gpu: info", g["info"], "iters", g["iters"], "rel L2 diff vs oracle", np.linalg.norm(g["x"] - r["x"]) / np.linalg.norm(r["x"]), flush=True)

import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
from oracle import oracle as O
from elmerfem_b200 import synth
A, b = synth.workload("heat", 200)
O.set_threads(O.max_threads())
ilu = O.ilu0(A)
xs = {}
for mode in (1, 2):
    O.set_dot_order(mode)
    t = time.time(); r = O.itersolve(A, b, method="bicgstab", precond="ilu0", ilu=ilu, tol=1e-8, maxit=2000)
    xs[mode] = r["x"]
    print("oracle C2 full size, dot order", mode, ": info", r["info"], "iters", r["iters"], "time %.1f s" % (time.time() - t), flush=True)
O.set_dot_order(0)
print("rel diff between orders 1 and 2:", np.linalg.norm(xs[1] - xs[2]) / np.linalg.norm(xs[1]))

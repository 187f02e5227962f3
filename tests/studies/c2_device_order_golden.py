"""Generates tests/golden/c2_device_order.json: the oracle's BiCGStab+ILU0 solve of BASELINE configs[1] (heat 200^3, tol 1e-8)
with its dot products summed in the DEVICE's order (oracle.set_dot_order(3), elmer_oracle.cpp dot_device): iteration count and the
SHA-256 of the solution's bytes.  A single-rank device solve must reproduce both (tests/test_gpu_bitwise.py).  CPU only.
  python tests/studies/c2_device_order_golden.py [orders...]     (default: 3 0 1 2)"""
import hashlib, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as O
from elmerfem_b200 import synth

orders = [int(a) for a in sys.argv[1:]] or [3, 0, 1, 2]
A, b = synth.workload("heat", 200)
O.set_threads(O.max_threads())
ilu = O.ilu0(A)
out_path = os.path.join(ROOT, "tests", "golden", "c2_device_order.json")
res = json.load(open(out_path)) if os.path.exists(out_path) else {}
res.update({"workload": "heat 200^3 hex8, default scaling, BiCGStab + ILU0, tol 1e-8", "n": int(A.n), "nnz": int(A.nnz),
            "ilu_sha256": hashlib.sha256(ilu.tobytes()).hexdigest(), "device_blocks": 148 * 8})
names = {0: "reference ddot (mathlibs, sequential)", 1: "eight interleaved partial sums", 2: "pairwise", 3: "device order"}
for mode in orders:
    O.set_dot_order(mode); O.set_device_blocks(148 * 8)
    t = time.time()
    r = O.itersolve(A, b, method="bicgstab", precond="ilu0", ilu=ilu, tol=1e-8, maxit=2000)
    res["order_%d" % mode] = {"what": names[mode], "info": r["info"], "iters": r["iters"], "residual": r["residual"],
                              "x_sha256": hashlib.sha256(r["x"].tobytes()).hexdigest(), "seconds": round(time.time() - t, 1)}
    print(mode, res["order_%d" % mode], flush=True)
    json.dump(res, open(out_path, "w"), indent=1)
O.set_dot_order(0)

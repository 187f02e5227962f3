"""Triangular-solve kernels side by side: bitwise comparison against the first variant + timing.
    python profiles/tools/sptrsv_modes.py 200 heat 0 3 4 -2            level kernel, wave tiles, lane tiles, the default (autotune)
    python profiles/tools/sptrsv_modes.py 200 heat 0 4,B200_LANE_E=3   level kernel vs lane tiles with request lead 3
A variant is `<B200_TRI_MODE>[,ENV=value,...]` (environment of earlier variants persists)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import elmerfem_b200 as B
from elmerfem_b200 import synth

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 200
kind = sys.argv[2] if len(sys.argv) > 2 else "heat"
variants = sys.argv[3:] or ["0", "1"]
A, b = synth.workload(kind, ne)
print("problem", kind, ne, "n", A.n, "nnz", A.nnz, flush=True)
v = np.random.RandomState(1).standard_normal(A.n)
ref = None
for var in variants:
    env = dict(kv.split("=") for kv in var.split(",") if "=" in kv)
    mode = var.split(",")[0]
    os.environ["B200_TRI_MODE"] = mode
    os.environ.update(env)
    M = B.Matrix()
    M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg)
    M.set_values(A.vals)
    t = time.time(); M.factorize(); tf = time.time() - t
    u = M.lu_precondition(v)
    if ref is None:
        ref = u
    same = np.array_equal(u, ref)
    ms = M.time_lu(5)
    st = M.stats()
    print("variant %-40s factor(wall) %.2fs factor_ms %.1f  lu %.3f ms  bitwise_same=%s nan=%d" % (var, tf, st.get("factor_ms", -1), ms, same, int(np.isnan(u).sum())), flush=True)
    if mode != "0" and len(sys.argv) > 3 and "solve" in env:
        pass
    M.close()

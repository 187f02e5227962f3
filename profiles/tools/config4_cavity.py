"""BASELINE configs[3]: linearised Navier-Stokes block system (4 dofs/node, nonsymmetric) on the unit cube, GCR and IDR(4) + ILU0,
resident solve on one GPU.  Full size: 135^3 elements = 136^3 nodes ~ 10.06 M dofs, ~1.1e9 nnz (13 GB CRS); run under gpurun:
    python profiles/tools/config4_cavity.py 135 gpurun_out/r02_c4.json
A smaller edge count (e.g. 96 -> 3.65 M dofs) fits a short call."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import elmerfem_b200 as B
from elmerfem_b200 import synth
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 135
t0 = time.time()
A, b = synth.workload("cavity", ne)
print("cavity %d^3: %d dofs, %d nnz, generated in %.1f s" % (ne, A.n, A.nnz, time.time() - t0), flush=True)
M = B.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg); M.set_values(A.vals)
M.factorize()
lv = M.levels()
print("ILU0 factor %.1f ms, levels %d forward / %d backward, SpMV %.3f ms, ILU0 application %.3f ms" %
      (M.stats()["factor_ms"], lv["forward"], lv["backward"], M.time_matvec(20), M.time_lu(10)), flush=True)
bytes_spmv = 12.0 * A.nnz + 20.0 * A.n + 4
bytes_lu = 12.0 * A.nnz + 4 * (A.n + 1) + 4 * A.n + 24 * A.n
peak = 6543.7
pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
spmv_ms, lu_ms = M.time_matvec(20), M.time_lu(10)
out = {"config": "C4: linearised Navier-Stokes block system (4 dofs/node, nonsymmetric), cavity %d^3 hex8, %d dofs, %d nnz, ILU0, tol 1e-8, 1 GPU" % (ne, A.n, A.nnz),
       "factor_ms": M.stats()["factor_ms"], "levels": [lv["forward"], lv["backward"]], "tri_mode": M.stats()["tri_mode"],
       "spmv": {"ms": spmv_ms, "gbs": bytes_spmv / spmv_ms / 1e6, "frac_of_hbm_peak": bytes_spmv / spmv_ms / 1e6 / peak, "bytes": bytes_spmv},
       "lu": {"ms": lu_ms, "gbs": bytes_lu / lu_ms / 1e6, "frac_of_hbm_peak": bytes_lu / lu_ms / 1e6 / peak, "bytes": bytes_lu}, "peak_gbs": peak, "solves": {}}
for method, kw in (("gcr", dict()), ("idrs", dict(idrs_s=4)), ("bicgstabl", dict(bicgstabl_l=4))):
    for rep in range(2):
        g = M.solve(b, method=method, precond="ilu0", tol=1e-8, maxit=2000, **kw)
    st = g["stats"]
    print("cavity %d^3 %-10s ilu0 iters %4d info %d solve %.1f ms -> %.1f it/s, %d launches" %
          (ne, method, g["iters"], g["info"], st["solve_ms"], g["iters"] / st["solve_ms"] * 1e3, st["launches"]), flush=True)
    r = A.to_scipy() @ g["x"] - b
    out["solves"][method] = {"iterations": g["iters"], "info": g["info"], "solve_ms": st["solve_ms"], "iterations_per_s": g["iters"] / st["solve_ms"] * 1e3,
                             "launches": st["launches"], "true_residual": float(np.linalg.norm(r) / np.linalg.norm(b))}
print("SpMV %.0f GB/s (CRS-equivalent bytes %.2f GB)" % (bytes_spmv / M.time_matvec(20) / 1e6, bytes_spmv / 1e9))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        json.dump(out, f, indent=1)
print(json.dumps(out))
M.close()

"""BASELINE configs[3]: linearised Navier-Stokes block system (4 dofs/node, nonsymmetric) on the unit cube, GCR and IDR(4) + ILU0,
resident solve on one GPU.  Full size: 135^3 elements = 136^3 nodes ~ 10.06 M dofs, ~1.1e9 nnz (13 GB CRS) -- not measured in
round 1 (GPU budget); run under gpurun:   python profiles/tools/config4_cavity.py 135
A smaller edge count (e.g. 96 -> 3.65 M dofs) fits a short call."""
import os, sys, time
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
for method, kw in (("gcr", dict()), ("idrs", dict(idrs_s=4)), ("bicgstabl", dict(bicgstabl_l=4))):
    for rep in range(2):
        g = M.solve(b, method=method, precond="ilu0", tol=1e-8, maxit=2000, **kw)
    st = g["stats"]
    print("cavity %d^3 %-10s ilu0 iters %4d info %d solve %.1f ms -> %.1f it/s, %d launches" %
          (ne, method, g["iters"], g["info"], st["solve_ms"], g["iters"] / st["solve_ms"] * 1e3, st["launches"]), flush=True)
print("SpMV %.0f GB/s (CRS-equivalent bytes %.2f GB)" % (bytes_spmv / M.time_matvec(20) / 1e6, bytes_spmv / 1e9))
M.close()

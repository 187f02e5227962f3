"""Lane-tile triangular solve (csrc/lane.cu) under the per-tile trace: time per application and where the tiles spend it.
    python profiles/tools/lane_lab.py 200 200 200 [ENV=value ...]      (elements per direction)
Prints: ms per application, per sweep the span (first start -> last end), the step time of tiles that never polled, the share of a
tile's life spent waiting for the ring / polling, and the finish time along the tile order."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
dims = tuple(int(a) for a in sys.argv[1:4])
for kv in sys.argv[4:]:
    k, v = kv.split("="); os.environ[k] = v
os.environ.setdefault("B200_TRI_MODE", "4")
trace = os.environ.get("LANE_LAB_TRACE", os.path.join(tempfile.gettempdir(), "lane_trace.txt"))
os.environ["B200_LANE_TRACE"] = trace
import elmerfem_b200 as B
from elmerfem_b200 import synth
A, b = synth.heat_cube(0, faces=["x0"], dims=dims)
v = np.random.RandomState(1).standard_normal(A.n)
M = B.Matrix(); M.set_structure(A.rows, A.cols, A.diag, 1, 1); M.set_values(A.vals); M.factorize()
for _ in range(4):
    u = M.lu_precondition(v)
ms = M.time_lu(5)
print("dims", dims, "n", A.n, " ".join(sys.argv[4:]), " lu %.3f ms per application, tri_mode %d" % (ms, M.stats()["tri_mode"]), flush=True)
M.close()
STEPS = (dims[0] + 1 + 62 + 2 * int(os.environ.get("B200_LANE_TC", "1")) + 7) // 8 * 8
if os.path.exists(trace):
    T = np.loadtxt(trace, ndmin=2)
    for sw in (0, 1):
        t = T[T[:, 0] == sw]
        if not len(t): continue
        dur = t[:, 5] - t[:, 4]
        span = t[:, 5].max() - t[:, 4].min()
        nt = len(t)
        free = t[t[:, 6] == 0]
        print(" sweep %d: %d tiles, span %.1f us, tile life median %.1f us (min %.1f, max %.1f)" % (sw, nt, span / 1e3, np.median(dur) / 1e3, dur.min() / 1e3, dur.max() / 1e3))
        if len(free):
            print("   tiles that never polled: %d, life %.1f us" % (len(free), np.median(free[:, 5] - free[:, 4]) / 1e3))
        run = t[:, 5] - np.where(t[:, 10] > 0, t[:, 10], t[:, 4])          # from the first step's operands to the end
        print("   first step ready -> end: median %.1f us (min %.1f max %.1f); ring wait %.0f kcyc, poll %.0f kcyc per tile (median), polls %.0f" %
              (np.median(run) / 1e3, run.min() / 1e3, run.max() / 1e3, np.median(t[:, 8]) / 1e3, np.median(t[:, 9]) / 1e3, np.median(t[:, 6])))
        if T.shape[1] >= 16:
            print("   cycles per tile (median) in: shuffles %.0fk, requests + ring test %.0fk, rows + stores %.0fk, ring -> registers %.0fk, replay resolve + update %.0fk" %
                  tuple(np.median(t[:, 11 + q]) / 1e3 for q in range(5)))
            for i in (0, nt // 2):
                print("   tile %d: life %.1f us, cycles/step: shuffles %.0f requests %.0f rows %.0f entry loads %.0f resolve %.0f (polls %d)" %
                      ((t[i, 1], (t[i, 5] - t[i, 4]) / 1e3) + tuple(t[i, 11 + q] / STEPS for q in range(5)) + (t[i, 6],)))
        idx = np.linspace(0, nt - 1, 12).astype(int)
        print("   tile: start/first/end us:", " ".join("%d:%.0f/%.0f/%.0f" % (t[i, 1], t[i, 4] / 1e3, t[i, 10] / 1e3, t[i, 5] / 1e3) for i in idx))

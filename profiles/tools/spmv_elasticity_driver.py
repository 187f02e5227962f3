"""ncu driver: a few SpMV on the 3-dof elasticity operand (block-column mode)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import elmerfem_b200 as B
from elmerfem_b200 import synth
p = synth.elasticity_slab(137, 137, 139, 0, 1)
n = p["rows"].size - 1
rowid = np.repeat(np.arange(1, n + 1, dtype=np.int64), np.diff(p["rows"]))
diag = (np.flatnonzero(p["cols"] == rowid) + 1).astype(np.int32)
M = B.Matrix(); M.set_structure(p["rows"], p["cols"], diag, 1, 3); M.set_values(p["vals"])
print("spmv ms", M.time_matvec(3))
M.close()

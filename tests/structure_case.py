"""Shared by test_structure.py (CPU) and test_zz_structure_gpu.py: a heat problem on a hex8 beam numbered the way
CreateMatrix numbers it (`Optimize Bandwidth = True`, the reference's default, fem/src/MainUtils.F90:1558-1560),
with the structure coming from the library's host-only producer (b200_node_graph / b200_optimize_bandwidth /
b200_initialize_structure)."""
import numpy as np

import elmerfem_b200 as b200
from elmerfem_b200 import synth

_cache = {}


def beam_heat_in_elmer_order(ex=24, ey=4, ez=4):
    key = (ex, ey, ez)
    if key in _cache:
        return _cache[key]
    xyz, elems = synth.grid_hex8(ex, ey, ez, 6.0, 1.0, 1.0)
    nn = xyz.shape[0]
    ptr = np.arange(0, 8 * elems.shape[0] + 1, 8, dtype=np.int32)
    S = b200.create_matrix_structure(ptr, elems.reshape(-1), nn, dofs=1)
    perm = S["perm"]
    _, hb_nat = b200.optimize_bandwidth(S["list_rows"], S["list_cols"], np.arange(1, nn + 1, dtype=np.int32), optimize=False)
    el_mat = np.ascontiguousarray(perm[elems - 1], dtype=np.int32)          # elements in matrix numbering
    xyz_mat = np.empty_like(xyz); xyz_mat[perm - 1] = xyz
    vals, rhs = synth.assemble(0, [1.0], xyz_mat, el_mat, 1, S["rows"], S["cols"], uniform=False)
    A = synth.CRS(S["rows"], S["cols"], S["diag"], vals, 1)
    fixed_nat = synth.boundary_nodes(ex, ey, ez, ["x0"])
    synth.dirichlet(A, rhs, np.sort(perm[fixed_nat - 1]).astype(np.int32), 0.0, False)
    # the same problem in ElmerGrid's own numbering
    r, c, d = synth.crs_structure(nn, elems, 1)
    v, bn = synth.assemble(0, [1.0], xyz, elems, 1, r, c, uniform=False)
    An = synth.CRS(r, c, d, v, 1)
    synth.dirichlet(An, bn, fixed_nat, 0.0, False)
    out = dict(A=A, b=rhs, perm=perm, nn=nn, half_bandwidth=S["half_bandwidth"], half_bandwidth_natural=hb_nat,
               accepted=not np.array_equal(perm, np.arange(1, nn + 1)), elems_matrix_numbering=el_mat,
               A_natural=An, b_natural=bn)
    _cache[key] = out
    return out

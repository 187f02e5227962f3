"""CPU: the synthetic generator numbers meshes exactly as the reference's ElmerGrid does, and the
mesh / partition readers reproduce the ownership rules of the MPI path (bit-exact integer work).

Fixtures under tests/golden/elmergrid/ were written by the reference's own ElmerGrid
(tests/golden/make_elmergrid_fixtures.sh): cube5.grd = 5 x 4 x 3 hex8 elements, and the same mesh
partitioned with `-partdual -metiskway 3`."""
import os

import numpy as np
import scipy.sparse as sp

from elmerfem_b200 import meshio, synth

EG = os.path.join(os.path.dirname(__file__), "golden", "elmergrid")


def test_grid_matches_elmergrid():
    hdr = meshio.read_header(EG)
    m = meshio.read_mesh(EG)
    assert hdr["nodes"] == 120 and hdr["elements"] == 60 and hdr["types"] == {404: 94, 808: 60}
    xyz, elems = synth.grid_hex8(5, 4, 3)
    assert np.array_equal(m.node_ids, np.arange(1, 121))
    assert np.allclose(m.xyz, xyz, atol=1e-12)                  # x-fastest numbering (Numbering = Horizontal)
    assert np.array_equal(np.array(m.elems), elems)             # connectivity, bit for bit
    # the boundary faces ElmerGrid wrote touch exactly the nodes our boundary selector returns
    bn = np.unique(np.concatenate([b[5] for b in m.bnd]))
    assert np.array_equal(bn, np.sort(synth.boundary_nodes(5, 4, 3, "all")))


def test_crs_structure_is_sorted_with_diag():
    xyz, elems = synth.grid_hex8(5, 4, 3)
    for ndof in (1, 3, 4):
        rows, cols, diag = synth.crs_structure(xyz.shape[0], elems, ndof)
        n = rows.size - 1
        assert n == 120 * ndof and rows[0] == 1 and rows[-1] - 1 == cols.size
        for i in range(n):
            seg = cols[rows[i] - 1:rows[i + 1] - 1]
            assert np.all(np.diff(seg) > 0)
            assert cols[diag[i] - 1] == i + 1
        # pattern = node adjacency through elements, expanded by ndof x ndof blocks
        adj = sp.lil_matrix((120, 120), dtype=np.int8)
        for e in elems:
            for a in e:
                adj[a - 1, e - 1] = 1
        assert cols.size == adj.nnz * ndof * ndof


def test_partition_ownership_and_numbering():
    P = meshio.Partitioning(os.path.join(EG, "partitioning.3"), 3)
    # every node is owned exactly once; owners come first in the neighbour lists of every sharer
    assert P.gn == 120 and P.goffset[0] == 0 and np.all(np.diff(P.goffset) > 0)
    seen = np.zeros(121, dtype=int)
    for p in range(3):
        ids = P.parts[p].node_ids
        for k, g in enumerate(ids):
            nl = P.shared[p].get(int(g))
            if nl is None:
                assert P.owner[p][k] == p
            else:
                assert p in nl and P.owner[p][k] == nl[0]
                for q in nl:                                   # all sharers agree on the list
                    assert P.shared[q][int(g)] == nl
        seen[ids[P.owner[p] == p]] += 1
    assert np.all(seen[1:] == 1)
    hdr = meshio.read_header(os.path.join(EG, "partitioning.3"), "part.1")
    assert hdr["shared"] == len(P.shared[0])
    # continuous numbering: a bijection, rank r's owned nodes are contiguous and keep their local order
    perm = P.dof_permutation()
    assert np.array_equal(np.sort(perm), np.arange(120))
    for p in range(3):
        mine = P.parts[p].node_ids[P.owner[p] == p]
        assert np.array_equal(perm[mine - 1], np.arange(P.goffset[p], P.goffset[p + 1]))


def test_heat_slab_equals_global_rows():
    ex, ey, ez, N = 6, 5, 13, 3
    xyz, el = synth.grid_hex8(ex, ey, ez, 1.0, ey / ex, ez / ex)
    r, c, d = synth.crs_structure(xyz.shape[0], el, 1)
    v, rhs = synth.assemble(0, [1.0], xyz, el, 1, r, c, uniform=True)
    A = synth.CRS(r, c, d, v, 1)
    synth.dirichlet(A, rhs, synth.boundary_nodes(ex, ey, ez, "all"), 0.0, False)
    synth.scale_system(A, rhs)
    S = A.to_scipy()
    parts = [synth.heat_slab(ex, ey, ez, k, N) for k in range(N)]
    tot = sum(float((p["b"] * p["bnorm"]) @ (p["b"] * p["bnorm"])) for p in parts)
    parts = [synth.heat_slab(ex, ey, ez, k, N, allreduce_sum=lambda s: tot) for k in range(N)]
    assert parts[0]["gn"] == A.n
    for k, p in enumerate(parts):
        lo, hi = p["goffset"][k], p["goffset"][k + 1]
        M = sp.csr_matrix((p["vals"], p["cols"] - 1, p["rows"] - 1), shape=(hi - lo, p["gn"]))
        G = S[lo:hi]
        assert np.array_equal(M.indptr, G.indptr) and np.array_equal(M.indices, G.indices)     # integers: bit-exact
        assert abs(M - G).max() < 1e-15 and np.abs(p["b"] - rhs[lo:hi]).max() < 1e-15


def test_linsys_dump_roundtrip(tmp_path):
    A, b = synth.heat_cube(3, faces=["x0"])
    S = A.to_scipy()
    meshio.write_linsys(S, b, dirname=str(tmp_path))
    S2, b2 = meshio.read_linsys(dirname=str(tmp_path))
    assert np.array_equal(S2.indptr, S.indptr) and np.array_equal(S2.indices, S.indices)
    assert np.array_equal(S2.data, S.data) and np.array_equal(b2, b)

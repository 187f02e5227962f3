"""Matrix-structure producer, CPU only: the library's host-only entry points b200_node_graph /
b200_optimize_bandwidth / b200_initialize_structure against the literal restatement of MakeListMatrix /
OptimizeBandwidth / InitializeMatrix in oracle/structure_oracle.cpp, plus independent constructions
(scipy patterns, numpy bandwidths).  Integer work: everything must be identical, not close."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import elmerfem_b200 as b200
from elmerfem_b200 import meshio, synth
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def _flat(elems):
    """list of node arrays (1-based) -> (elem_ptr 0-based, elem_nodes)"""
    ptr = np.zeros(len(elems) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(e) for e in elems])
    return ptr, np.concatenate([np.asarray(e, dtype=np.int32) for e in elems]).astype(np.int32)


def _hex(ex, ey, ez):
    _, elems = synth.grid_hex8(ex, ey, ez)
    nn = (ex + 1) * (ey + 1) * (ez + 1)
    return nn, [e for e in elems]


def _graph_from_scipy(S):
    S = sp.csr_matrix(S); S.sort_indices()
    return (S.indptr + 1).astype(np.int32), (S.indices + 1).astype(np.int32)


def _random_graph(rng, n, avg_deg, ncomp=1):
    """symmetric pattern with full diagonal, ncomp disconnected blocks"""
    blocks = []
    sizes = np.full(ncomp, n // ncomp); sizes[: n % ncomp] += 1
    for m in sizes:
        R = sp.random(m, m, density=min(1.0, avg_deg / max(m, 1)), random_state=rng.integers(1 << 31), format="csr")
        blocks.append(((R + R.T) != 0).astype(np.int8) + sp.identity(m, dtype=np.int8, format="csr"))
    S = sp.block_diag(blocks, format="csr")
    p = rng.permutation(n)  # hide the block structure
    return _graph_from_scipy(S[p][:, p])


def _half_bandwidth(rows, cols, number=None):
    n = rows.size - 1
    r = np.repeat(np.arange(n), np.diff(rows)); c = cols - 1
    if number is not None:
        r, c = number[r], number[c]
    return int(np.abs(r - c).max(initial=0)) + 1


def _elmergrid_meshes():
    base = os.path.join(HERE, "golden", "elmergrid")
    return [d for d in [base] + [os.path.join(base, x) for x in sorted(os.listdir(base))]
            if os.path.exists(os.path.join(d, "mesh.elements"))]


def _mesh_elements(dirname):
    """bulk then boundary elements with node numbers compacted to 1..nn, as Elmer's mesh reader stores them"""
    m = meshio.read_mesh(dirname)
    nid = np.zeros(int(m.node_ids.max()) + 1, dtype=np.int64)
    nid[m.node_ids] = np.arange(1, m.node_ids.size + 1)
    elems = [nid[c] for c in m.elems] + [nid[b[5]] for b in m.bnd]
    return m.node_ids.size, elems


# ------------------------------------------------------------------------------------------ node graph

@pytest.mark.parametrize("dims", [(1, 1, 1), (3, 2, 4), (6, 6, 6)])
def test_node_graph_hex_three_ways(dims):
    nn, elems = _hex(*dims)
    ptr, nodes = _flat(elems)
    ident = np.arange(1, nn + 1, dtype=np.int32)
    r_o, c_o = orc.make_list_matrix(ptr, nodes, ident, nn)
    r_p, c_p = b200.node_graph(ptr, nodes, nn)                 # perm = NULL
    r_q, c_q = b200.node_graph(ptr, nodes, nn, ident, nn)
    r_s, c_s, _ = synth.crs_structure(nn, np.array(elems, dtype=np.int32), 1)   # third, independent construction
    for r, c in ((r_p, c_p), (r_q, c_q), (r_s, c_s)):
        assert np.array_equal(r, r_o) and np.array_equal(c, c_o)
    # 0-based variant: same graph shifted
    r_z, c_z = b200.node_graph(ptr, nodes - 1, nn, index_base=0)
    assert np.array_equal(r_z + 1, r_o) and np.array_equal(c_z + 1, c_o)


def test_node_graph_partial_equation_and_mixed_elements():
    rng = np.random.default_rng(7)
    nn = 60
    elems = [rng.choice(nn, size=rng.integers(1, 9), replace=False) + 1 for _ in range(80)]   # 101..808-like sizes
    ptr, nodes = _flat(elems)
    active = rng.random(nn) < 0.7
    perm = np.zeros(nn, dtype=np.int32)
    order = rng.permutation(np.nonzero(active)[0])
    perm[order] = np.arange(1, order.size + 1)
    k = int(order.size)
    r_o, c_o = orc.make_list_matrix(ptr, nodes, perm, k)
    r_p, c_p = b200.node_graph(ptr, nodes, nn, perm, k)
    assert np.array_equal(r_p, r_o) and np.array_equal(c_p, c_o)
    assert all(np.all(np.diff(c_p[r_p[i] - 1:r_p[i + 1] - 1]) > 0) for i in range(k))


def test_node_graph_rejects_bad_input():
    ptr, nodes = _flat([[1, 2, 9]])
    with pytest.raises(b200.B200Error):
        b200.node_graph(ptr, nodes, 3)


# ------------------------------------------------------------------------------------------ OptimizeBandwidth

def _check_optimize(rows, cols, perm, use_optimized):
    p_o, hb_o = orc.optimize_bandwidth(rows, cols, perm, True, use_optimized)
    p_p, hb_p = b200.optimize_bandwidth(rows, cols, perm, True, use_optimized)
    assert hb_p == hb_o
    assert np.array_equal(p_p, p_o)
    # independent checks of what came back
    k = rows.size - 1
    act = np.asarray(perm) > 0
    assert np.array_equal(np.sort(p_p[act]), np.arange(1, k + 1))
    assert np.all(p_p[~act] == 0) or np.array_equal(p_p, perm)
    number = np.zeros(k, dtype=np.int64)
    number[np.asarray(perm)[act] - 1] = p_p[act] - 1
    assert hb_p == _half_bandwidth(rows, cols, number)
    before = _half_bandwidth(rows, cols)
    if not use_optimized:
        assert hb_p <= before
        if np.array_equal(p_p, perm):
            assert hb_p == before
    return p_p, hb_p


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (5, 4, 3), (3, 9, 2), (10, 10, 10)])
@pytest.mark.parametrize("use_optimized", [False, True])
def test_optimize_bandwidth_hex(dims, use_optimized):
    nn, elems = _hex(*dims)
    ptr, nodes = _flat(elems)
    rows, cols = b200.node_graph(ptr, nodes, nn)
    _check_optimize(rows, cols, np.arange(1, nn + 1, dtype=np.int32), use_optimized)


@pytest.mark.parametrize("seed", range(12))
def test_optimize_bandwidth_random_graphs(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 400))
    rows, cols = _random_graph(rng, n, avg_deg=float(rng.uniform(0.5, 6.0)), ncomp=int(rng.integers(1, 5)))
    # Perm longer than k with inactive entries, active ones in random initial order
    m = n + int(rng.integers(0, 20))
    perm = np.zeros(m, dtype=np.int32)
    perm[rng.permutation(m)[:n]] = rng.permutation(n) + 1
    for use_optimized in (False, True):
        _check_optimize(rows, cols, perm, use_optimized)


def test_start_node_search_never_moves():
    """The start-node refinement loop (BandwidthOptimize.F90:247-271) only moves to a node of strictly lower degree
    than the current start, and the initial start already has the globally lowest degree (222-229): the loop body,
    including `StartNode = j` at :268, cannot run.  The oracle counts how often it did; the sweep therefore always
    starts at the first node of minimum degree, which is checked through the number that node receives (k, since the
    numbering is reversed)."""
    for seed in range(60):
        rng = np.random.default_rng(1000 + seed)
        n = int(rng.integers(5, 80))
        rows, cols = _random_graph(rng, n, avg_deg=float(rng.uniform(1.0, 3.0)), ncomp=int(rng.integers(1, 3)))
        perm = np.arange(1, n + 1, dtype=np.int32)
        p_o, _ = orc.optimize_bandwidth(rows, cols, perm, True, True)
        assert orc.lib().orc_optimize_bandwidth_new_roots() == 0
        p_p, _ = b200.optimize_bandwidth(rows, cols, perm, True, True)
        first_min = int(np.argmin(np.diff(rows)))
        assert p_p[first_min] == n and p_o[first_min] == n


def test_optimize_bandwidth_off_reports_initial_bandwidth():
    nn, elems = _hex(4, 3, 2)
    ptr, nodes = _flat(elems)
    rows, cols = b200.node_graph(ptr, nodes, nn)
    perm = np.arange(1, nn + 1, dtype=np.int32)
    p_o, hb_o = orc.optimize_bandwidth(rows, cols, perm, False, False)
    p_p, hb_p = b200.optimize_bandwidth(rows, cols, perm, False, False)
    assert hb_p == hb_o == _half_bandwidth(rows, cols)
    assert np.array_equal(p_p, perm) and np.array_equal(p_o, perm)


def test_optimize_bandwidth_shuffled_mesh_is_improved():
    """A scrambled node numbering of a beam must come back with a much smaller bandwidth (accepted ordering)."""
    nn, elems = _hex(24, 3, 3)
    rng = np.random.default_rng(3)
    shuffle = rng.permutation(nn) + 1
    elems = [shuffle[e - 1] for e in elems]
    ptr, nodes = _flat(elems)
    rows, cols = b200.node_graph(ptr, nodes, nn)
    perm = np.arange(1, nn + 1, dtype=np.int32)
    p, hb = _check_optimize(rows, cols, perm, False)
    assert hb < _half_bandwidth(rows, cols) // 4
    assert not np.array_equal(p, perm)


@pytest.mark.parametrize("meshdir", _elmergrid_meshes())
def test_optimize_bandwidth_elmergrid_fixtures(meshdir):
    nn, elems = _mesh_elements(meshdir)
    ptr, nodes = _flat(elems)
    rows, cols = b200.node_graph(ptr, nodes, nn)
    r_o, c_o = orc.make_list_matrix(ptr, nodes, np.arange(1, nn + 1, dtype=np.int32), nn)
    assert np.array_equal(rows, r_o) and np.array_equal(cols, c_o)
    _check_optimize(rows, cols, np.arange(1, nn + 1, dtype=np.int32), False)
    _check_optimize(rows, cols, np.arange(1, nn + 1, dtype=np.int32), True)


def test_optimize_bandwidth_winkel_mesh():
    import winkel_case as wk
    if not wk.available():
        pytest.skip("oracle/_ref/ElmerGrid not built")
    nn, elems = _mesh_elements(wk.mesh_dir())
    ptr, nodes = _flat(elems)
    rows, cols = b200.node_graph(ptr, nodes, nn)
    _check_optimize(rows, cols, np.arange(1, nn + 1, dtype=np.int32), False)


# ------------------------------------------------------------------------------------------ InitializeMatrix

@pytest.mark.parametrize("dofs", [1, 3])
@pytest.mark.parametrize("reorder", [False, True])
def test_initialize_structure(dofs, reorder):
    nn, elems = _hex(4, 3, 5)
    ptr, nodes = _flat(elems)
    rows, cols = b200.node_graph(ptr, nodes, nn)
    perm0 = np.arange(1, nn + 1, dtype=np.int32)
    if reorder:
        perm1, _ = b200.optimize_bandwidth(rows, cols, perm0, True, True)
        got = b200.initialize_structure(rows, cols, dofs, perm0, perm1)
        ref = orc.initialize_matrix(rows, cols, dofs, perm0, perm1)
        number = perm1 - 1
    else:
        got = b200.initialize_structure(rows, cols, dofs)
        ref = orc.initialize_matrix(rows, cols, dofs)
        number = np.arange(nn)
    for g, r in zip(got, ref):
        assert np.array_equal(g, r)
    # independent: permuted pattern (x) dense dofs x dofs block
    G = sp.csr_matrix((np.ones(cols.size), cols - 1, rows - 1), shape=(nn, nn)).tocoo()
    P = sp.csr_matrix((np.ones(G.nnz), (number[G.row], number[G.col])), shape=(nn, nn))
    K = sp.kron(P, np.ones((dofs, dofs)), format="csr"); K.sort_indices()
    R, Cc, D = got
    assert np.array_equal(R - 1, K.indptr) and np.array_equal(Cc - 1, K.indices)
    assert np.array_equal(Cc[D - 1], np.arange(1, nn * dofs + 1))
    # 0-based call gives the same arrays shifted
    if reorder:
        z = b200.initialize_structure(rows - 1, cols - 1, dofs, perm0, perm1, index_base=0)
    else:
        z = b200.initialize_structure(rows - 1, cols - 1, dofs, index_base=0)
    for g, zz in zip(got, z):
        assert np.array_equal(g - 1, zz)


def test_create_matrix_structure_beam_solution_is_the_permuted_one(oracle):
    """CreateMatrix's nodal path end to end on a beam, where the optimiser's numbering is accepted (the cube keeps
    its ElmerGrid numbering, see test_optimize_bandwidth_hex): structure from the library, values assembled on it,
    oracle BiCGStab+ILU0.  The solution must be the natural-order solution permuted."""
    import structure_case as sc
    S = sc.beam_heat_in_elmer_order()
    assert S["accepted"] and S["half_bandwidth"] < S["half_bandwidth_natural"]
    A, b, perm = S["A"], S["b"], S["perm"]
    # structure the library produced == structure built directly on the renumbered elements
    r2, c2, d2 = synth.crs_structure(S["nn"], S["elems_matrix_numbering"], 1)
    assert np.array_equal(A.rows, r2) and np.array_equal(A.cols, c2) and np.array_equal(A.diag, d2)
    got = oracle.itersolve(A, b, method="bicgstab", precond="ilu0", tol=1e-10, maxit=500)
    An, bn = S["A_natural"], S["b_natural"]
    ref = oracle.itersolve(An, bn, method="bicgstab", precond="ilu0", tol=1e-10, maxit=500)
    assert got["info"] == 1 and ref["info"] == 1
    x_back = got["x"][perm - 1]                       # mesh node m is row perm[m]
    assert np.linalg.norm(x_back - ref["x"]) <= 1e-7 * np.linalg.norm(ref["x"])


def test_empty_and_degenerate_inputs():
    """No elements / no active nodes / a single node: sizes come back consistent and nothing is touched out of bounds."""
    ptr = np.zeros(1, dtype=np.int32)                                  # zero elements
    rows, cols = b200.node_graph(ptr, np.zeros(1, dtype=np.int32), 5, np.zeros(5, dtype=np.int32), 0)
    assert rows.tolist() == [1] and cols.size == 0
    p, hb = b200.optimize_bandwidth(rows, np.ones(1, dtype=np.int32), np.zeros(5, dtype=np.int32))
    assert hb == 1 and p.tolist() == [0] * 5
    R, Cc, D = b200.initialize_structure(rows, np.ones(1, dtype=np.int32), 2)
    assert R.tolist() == [1] and Cc.size == 0
    # one node, one 1-node element (a 101 point element)
    ptr, nodes = _flat([[1]])
    rows, cols = b200.node_graph(ptr, nodes, 1)
    assert rows.tolist() == [1, 2] and cols.tolist() == [1]
    p, hb = b200.optimize_bandwidth(rows, cols, np.array([1], dtype=np.int32))
    assert p.tolist() == [1] and hb == 1
    r_o, c_o = orc.make_list_matrix(ptr, nodes, np.array([1], dtype=np.int32), 1)
    p_o, hb_o = orc.optimize_bandwidth(r_o, c_o, np.array([1], dtype=np.int32))
    assert p_o.tolist() == [1] and hb_o == 1
    R, Cc, D = b200.initialize_structure(rows, cols, 3)
    assert R.tolist() == [1, 4, 7, 10] and Cc.tolist() == [1, 2, 3] * 3 and D.tolist() == [1, 5, 9]


def test_structure_entry_points_reject_inconsistent_input():
    rows = np.array([1, 3, 4], dtype=np.int32); cols = np.array([1, 2, 7], dtype=np.int32)      # column 7 of a 2-row graph
    with pytest.raises(b200.B200Error):
        b200.optimize_bandwidth(rows, cols, np.array([1, 2], dtype=np.int32))
    good = np.array([1, 2, 2], dtype=np.int32)
    with pytest.raises(b200.B200Error):
        b200.optimize_bandwidth(rows, good, np.array([1, 5], dtype=np.int32))                    # Perm entry beyond k
    with pytest.raises(b200.B200Error):
        b200.initialize_structure(rows, good, 0)                                                 # dofs < 1
    with pytest.raises(b200.B200Error):
        b200.initialize_structure(rows, good, 1, np.array([1, 2], dtype=np.int32), np.array([1, 0], dtype=np.int32))   # numbering misses a row

"""CPU: the host-side planning of the multi-GPU path (send lists, ghost slots, owned/ghost split) is
bit-exact against a restatement of the reference's elmer_distribute_matrix (fem/src/rocalution.cpp),
on z-slab partitions, on the reference's own METIS partition of an ElmerGrid mesh, and across two real
processes talking over gloo."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from elmerfem_b200 import meshio, synth

EG = os.path.join(os.path.dirname(__file__), "golden", "elmergrid")


def plan_all(b200, parts, goffset):
    """What b200_set_partition does on every rank, with the exchange done by list copies."""
    nr = len(parts)
    gn = int(goffset[-1])
    send = []
    for r, (rows, cols, _) in enumerate(parts):
        cnt, gid = b200.partition_send_lists(gn, rows, cols, goffset, r, index_base=0)
        ptr = np.concatenate([[0], np.cumsum(cnt)])
        send.append([gid[ptr[q]:ptr[q + 1]] for q in range(nr)])
    plans = []
    for r, (rows, cols, _) in enumerate(parts):
        neigh = [q for q in range(nr) if q != r and (len(send[r][q]) or len(send[q][r]))]
        send_idx = np.concatenate([send[r][q] - goffset[r] for q in neigh] + [np.zeros(0, dtype=np.int32)]).astype(np.int32)
        ghost = np.concatenate([send[q][r] for q in neigh] + [np.zeros(0, dtype=np.int32)]).astype(np.int32)
        sp_ = np.concatenate([[0], np.cumsum([len(send[r][q]) for q in neigh])]).astype(np.int32)
        rp_ = np.concatenate([[0], np.cumsum([len(send[q][r]) for q in neigh])]).astype(np.int32)
        split = b200.partition_split(rows, cols, int(goffset[r]), int(goffset[r + 1]), ghost, index_base=0)
        plans.append(dict(neigh=np.array(neigh, dtype=np.int32), send_ptr=sp_, send_idx=send_idx, recv_ptr=rp_, ghost_gid=ghost, **split))
    # peer-memory halo path: what a sender derives about a neighbour's receive area from the send-count matrix alone
    # must be that neighbour's actual plan (its neighbour order, its receive offsets)
    cnt = np.array([[len(send[s_][d]) for d in range(nr)] for s_ in range(nr)], dtype=np.int32)
    for me in range(nr):
        for r in range(nr):
            if r == me:
                continue
            q, nn, ng, off = b200.partition_peer_layout(nr, me, r, cnt)
            nb = plans[r]["neigh"].tolist()
            assert nn == len(nb) and ng == len(plans[r]["ghost_gid"])
            if me in nb:
                assert q == nb.index(me) and off == int(plans[r]["recv_ptr"][q])
                assert int(plans[r]["recv_ptr"][q + 1]) - off == cnt[me, r]
            else:
                assert q == -1
    return plans


def check_against_oracle(b200, S, goffset):
    from oracle import halo_oracle as HO
    parts = HO.split_rows(S, goffset)
    ref = HO.distribute([(p[0], p[1]) for p in parts], list(goffset))
    got = plan_all(b200, parts, np.asarray(goffset, dtype=np.int32))
    rng = np.random.RandomState(0)
    x = rng.standard_normal(S.shape[0])
    y = np.zeros_like(x)
    for r, (g, o) in enumerate(zip(got, ref)):
        for k in ["neigh", "send_ptr", "send_idx", "recv_ptr", "ghost_gid"]:
            assert np.array_equal(g[k], o[k]), (r, k)
        assert np.array_equal(g["oo_rows"], o["int_rows"]) and np.array_equal(g["oo_cols"], o["int_cols"])
        n = len(parts[r][0]) - 1
        grow = np.repeat(np.arange(n), np.diff(g["g_rows"]))
        assert np.array_equal(grow, o["ghost_row"]) and np.array_equal(g["g_cols"] - n, o["ghost_col"])
        # emulate the halo SpMV of this rank with the plan and compare with the global product
        lo, hi = goffset[r], goffset[r + 1]
        xl = np.concatenate([x[lo:hi], x[g["ghost_gid"]]])
        rows, cols, vals = parts[r]
        own = (cols >= lo) & (cols < hi)
        Aoo = sp.csr_matrix((vals[own], g["oo_cols"], g["oo_rows"]), shape=(n, n))
        Ag = sp.csr_matrix((vals[~own], g["g_cols"], g["g_rows"]), shape=(n, n + len(g["ghost_gid"])))
        y[lo:hi] = Aoo @ xl[:n] + Ag @ xl
        assert np.array_equal(np.sort(xl[n:]), np.sort(x[g["ghost_gid"]]))
    assert np.abs(y - S @ x).max() <= 1e-12 * np.abs(S @ x).max()
    return got


def test_slab_partition_plan(b200):
    A, _ = synth.heat_cube(0, faces="all", dims=(4, 3, 9))
    S = A.to_scipy()
    plane = 5 * 4
    goffset = [plane * l for l in synth.slab_layers(10, 3)]
    got = check_against_oracle(b200, S, goffset)
    assert list(got[1]["neigh"]) == [0, 2] and list(got[0]["neigh"]) == [1]


def test_metis_partition_plan(b200):
    """The reference's METIS k-way partition of the ElmerGrid mesh -> ownership -> continuous numbering ->
    complete owned rows -> halo plan."""
    P = meshio.Partitioning(os.path.join(EG, "partitioning.3"), 3, ndof=1)
    A, _ = synth.heat_cube(0, faces=["x0"], dims=(5, 4, 3))
    parts, Sc = P.owned_rows(A.to_scipy())
    check_against_oracle(b200, Sc, P.goffset)
    # 3 dofs per node: dof ownership copied from the nodes (ParallelUtils.F90:180-190)
    P3 = meshio.Partitioning(os.path.join(EG, "partitioning.3"), 3, ndof=3)
    A3, _ = synth.elasticity_beam(5, 4, 3, lx=1.0)
    parts3, Sc3 = P3.owned_rows(A3.to_scipy())
    check_against_oracle(b200, Sc3, P3.goffset)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, out_q):
    import torch.distributed as dist
    import torch
    import elmerfem_b200 as B
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        ex, ey, ez = 5, 4, 11

        def allsum(v):
            t = torch.tensor([v], dtype=torch.float64); dist.all_reduce(t); return float(t.item())
        p = synth.heat_slab(ex, ey, ez, rank, world, allreduce_sum=allsum)
        goff = p["goffset"]
        cnt, gid = B.partition_send_lists(p["gn"], p["rows"], p["cols"], goff, rank, index_base=1)
        # exchange exactly as the library does with NCCL: counts (all-gather), then the lists
        allcnt = [None] * world
        dist.all_gather_object(allcnt, cnt.tolist())
        ptr = np.concatenate([[0], np.cumsum(cnt)])
        lists = [gid[ptr[q]:ptr[q + 1]].tolist() for q in range(world)]
        everything = [None] * world
        dist.all_gather_object(everything, lists)
        neigh = [q for q in range(world) if q != rank and (cnt[q] or allcnt[q][rank])]
        ghost = np.array([g for q in neigh for g in everything[q][rank]], dtype=np.int32)
        split = B.partition_split(p["rows"], p["cols"], int(goff[rank]), int(goff[rank + 1]), ghost, index_base=1)
        # distributed SpMV + dot with the plan: x_global = arange-based vector every rank can evaluate
        n = p["rows"].size - 1
        xg = lambda ids: np.sin(0.37 * np.asarray(ids, dtype=np.float64) + 1.0)     # noqa: E731
        own_ids = np.arange(goff[rank], goff[rank + 1])
        # halo exchange: send x at my send lists, receive ghosts
        xl = np.concatenate([xg(own_ids), np.zeros(len(ghost))])
        sendvals = [xl[np.array(everything[rank][q], dtype=np.int64) - goff[rank]].tolist() if q in neigh else [] for q in range(world)]
        allvals = [None] * world
        dist.all_gather_object(allvals, sendvals)
        k = n
        for q in neigh:
            v = allvals[q][rank]
            xl[k:k + len(v)] = v
            k += len(v)
        own = (p["cols"] - 1 >= goff[rank]) & (p["cols"] - 1 < goff[rank + 1])
        Aoo = sp.csr_matrix((p["vals"][own], split["oo_cols"], split["oo_rows"]), shape=(n, n))
        Ag = sp.csr_matrix((p["vals"][~own], split["g_cols"], split["g_rows"]), shape=(n, n + len(ghost)))
        y = Aoo @ xl[:n] + Ag @ xl
        # reference: the same rows applied to the globally evaluated x
        Aglob = sp.csr_matrix((p["vals"], p["cols"] - 1, p["rows"] - 1), shape=(n, p["gn"]))
        yref = Aglob @ xg(np.arange(p["gn"]))
        dot = allsum(float(y @ y)); dref = allsum(float(yref @ yref))
        ok = bool(np.abs(y - yref).max() <= 1e-13 * max(1.0, np.abs(yref).max())) and abs(dot - dref) <= 1e-12 * dref
        ok = ok and np.array_equal(xl[n:], xg(ghost))
        out_q.put((rank, ok, len(neigh), int(len(ghost))))
    finally:
        dist.destroy_process_group()


def test_two_process_gloo_halo_exchange(b200):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, nneigh, nghost in res:
        assert ok, rank
        assert nneigh == 1 and nghost == 6 * 5          # one interface plane of (ex+1)*(ey+1) nodes


def test_vectorised_restatement_equals_the_loop_restatement():
    """oracle/halo_oracle.py: send_lists_rank + plan_rank (numpy, used by bench.py at full size) against distribute() (the literal
    loops of rocalution.cpp:64-372) on random sparse matrices with uneven ownership ranges."""
    import scipy.sparse as sp
    from oracle import halo_oracle as HO
    rs = np.random.RandomState(5)
    for n, nparts in [(60, 2), (97, 3), (200, 5), (40, 8)]:
        S = sp.random(n, n, density=0.08, random_state=rs, format="csr")
        S = (S + S.T + sp.identity(n, format="csr")).tocsr()          # structurally symmetric, as finite-element matrices are (the reference relies on it)
        cuts = np.sort(rs.choice(np.arange(1, n), nparts - 1, replace=False))
        off = [0] + [int(c) for c in cuts] + [n]
        parts = HO.split_rows(S, off)
        ref = HO.distribute([(p[0], p[1]) for p in parts], off)
        sends = [HO.send_lists_rank(p[0], p[1], off, r) for r, p in enumerate(parts)]
        for r in range(nparts):
            got = HO.plan_rank(sends, off, r)
            for k in ["neigh", "send_ptr", "send_idx", "recv_ptr", "ghost_gid"]:
                assert np.array_equal(got[k], ref[r][k]), (n, nparts, r, k)

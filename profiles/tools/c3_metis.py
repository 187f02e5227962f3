#!/usr/bin/env python
"""BASELINE configs[2] (C3) at its stated size on the reference's own partitions: linear elasticity (3 dofs per node) on a hex8 beam of
~25 M dofs, meshed AND partitioned by the reference's ElmerGrid (`ElmerGrid 1 2 beam -partdual -metiskway N`, oracle/_ref/ElmerGrid, built
from /root/reference/elmergrid/src), BiCGStab(l=4) + block-Jacobi ILU0 on N GPUs, one rank per partition.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/tools/c3_metis.py --ex 400 --ey 143 --ez 143 --out gpurun_out/r02_c3.json
    python profiles/tools/c3_metis.py --dry --nparts 4 --ex 24 --ey 8 --ez 8          (CPU: builds every rank's system, checks it against a global assembly)

Every rank reads the mesh and the partition files ElmerGrid wrote (mesh.nodes / mesh.elements, partitioning.N/part.k.nodes and .shared),
derives the dof ownership (owner = first entry of the shared list, SParIterSolver.F90:232) and the continuous numbering
(SParIterSolver.F90:1453-1488), assembles the elements that touch its owned nodes, keeps its complete owned rows with global continuous
column ids, applies Elmer's diagonal scaling (owners' diagonals exchanged), and hands the rows to b200_set_partition.  Halo index lists are
compared with the numpy restatement of elmer_distribute_matrix on every rank.  Not part of the product: a measurement driver."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
ELMERGRID = os.path.join(ROOT, "oracle", "_ref", "ElmerGrid")

GRD = """***** ElmerGrid input file for structured grid generation *****
Version = 210903
Coordinate System = Cartesian 3D
Subcell Divisions in 3D = 1 1 1
Subcell Sizes 1 = %g
Subcell Sizes 2 = %g
Subcell Sizes 3 = %g
Material Structure in 2D
  1
End
Materials Interval = 1 1
Boundary Definitions
! type     out      int
  1        0        1        1
End
Numbering = Horizontal
Element Degree = 1
Element Innernodes = False
Element Divisions 1 = %d
Element Divisions 2 = %d
Element Divisions 3 = %d
"""


def make_mesh(workdir, ex, ey, ez, nparts):
    """ElmerGrid: structured hex8 beam (cubic cells of side 1/ey) + METIS k-way partition of the dual graph."""
    h = 1.0 / ey
    with open(os.path.join(workdir, "beam.grd"), "w") as f:
        f.write(GRD % (ex * h, ey * h, ez * h, ex, ey, ez))
    t = time.time()
    cmd = [ELMERGRID, "1", "2", "beam"] + (["-partdual", "-metiskway", str(nparts)] if nparts > 1 else [])
    subprocess.check_call(cmd, cwd=workdir, stdout=open(os.path.join(workdir, "elmergrid.log"), "w"), stderr=subprocess.STDOUT)
    return time.time() - t


def read_table(path, ncols=None):
    import pandas as pd
    return pd.read_csv(path, sep=r"\s+", header=None, engine="c").to_numpy()


def ownership(meshdir, nparts):
    """owner[node-1] (0-based partition) and, per partition, its nodes in file order.  Owner = first entry of the node's line in
    part.k.shared (SParIterSolver.F90:232); nodes that are not shared belong to the partition whose file lists them."""
    files = []
    pdir = os.path.join(meshdir, "partitioning.%d" % nparts)
    for p in range(nparts):
        files.append(read_table(os.path.join(pdir, "part.%d.nodes" % (p + 1)))[:, 0].astype(np.int64))
    nn = int(max(f.max() for f in files))
    owner = np.full(nn, -1, dtype=np.int32)
    for p in range(nparts):
        ids = files[p]
        owner[ids - 1] = np.where(owner[ids - 1] < 0, p, owner[ids - 1])       # provisional: first partition that lists the node
    for p in range(nparts):
        sp = os.path.join(pdir, "part.%d.shared" % (p + 1))
        if not os.path.exists(sp):
            continue
        with open(sp) as f:
            for line in f:
                t = line.split()
                if len(t) >= 3:
                    owner[int(t[0]) - 1] = int(t[2]) - 1                       # node, count, owner, other sharers ...
    assert (owner >= 0).all()
    return owner, files


def rank_system(meshdir, nparts, rank, owner, files, xyz, elems, allgather_D=None, allsum=None, E=1e9, nu=0.3, load=(0.0, -1e4, 0.0)):
    """Complete owned rows of rank `rank` (1-based CRS, GLOBAL continuous column ids), scaled; + rhs, goffset, gn."""
    from elmerfem_b200 import synth
    import scipy.sparse as sp
    counts = np.array([int(np.count_nonzero(owner[files[p] - 1] == p)) for p in range(nparts)], dtype=np.int64)
    goff_nodes = np.concatenate([[0], np.cumsum(counts)])
    cont = np.full(owner.size, -1, dtype=np.int64)                               # node -> continuous node number
    for p in range(nparts):
        mine = files[p][owner[files[p] - 1] == p]
        cont[mine - 1] = goff_nodes[p] + np.arange(mine.size)
    assert (cont >= 0).all()
    own_nodes = files[rank][owner[files[rank] - 1] == rank]                      # file order = continuous order
    touch = (owner[elems - 1] == rank).any(axis=1)
    el = elems[touch]
    ghosts = np.setdiff1d(np.unique(el), own_nodes)
    lnodes = np.concatenate([own_nodes, ghosts])
    lid = np.zeros(owner.size + 1, dtype=np.int64)
    lid[lnodes] = np.arange(1, lnodes.size + 1)
    lel = np.ascontiguousarray(lid[el].astype(np.int32))
    lxyz = np.ascontiguousarray(xyz[lnodes - 1])
    rows, cols, diag = synth.crs_structure(lnodes.size, lel, 3)
    vals, rhs = synth.assemble(1, [E, nu, load[0], load[1], load[2]], lxyz, lel, 3, rows, cols, uniform=False)
    A = synth.CRS(rows, cols, diag, vals, 3)
    fixed = np.flatnonzero(lxyz[:, 0] <= 1e-12) + 1                              # clamped end x = 0 (all local nodes on it)
    if fixed.size:
        dofs = np.sort(np.concatenate([3 * (fixed - 1) + c + 1 for c in range(3)])).astype(np.int32)
        synth.dirichlet(A, rhs, dofs, 0.0, False)
    no = own_nodes.size * 3
    p1 = A.rows[no] - 1
    lrows0 = (A.rows[:no + 1] - 1).astype(np.int64)
    ldof_to_g = (np.repeat(cont[lnodes - 1] * 3, 3) + np.tile(np.arange(3), lnodes.size)).astype(np.int64)
    gcols0 = ldof_to_g[A.cols[:p1] - 1]
    d_own = np.abs(A.vals[A.diag[:no] - 1])
    D_own = 1.0 / np.sqrt(d_own)
    gn = int(goff_nodes[-1] * 3)
    if allgather_D is not None:
        Dg = allgather_D(D_own, (goff_nodes * 3).astype(np.int64))               # owners' scaling factors of every dof
    else:
        Dg = None
    S = sp.csr_matrix((A.vals[:p1], gcols0, lrows0), shape=(no, gn))
    S.sort_indices()                                                             # ascending global columns inside every row
    rowid = np.repeat(np.arange(no), np.diff(S.indptr))
    if Dg is not None:
        S.data *= D_own[rowid] * Dg[S.indices]
    b = rhs[:no] * D_own
    s = float(np.dot(b, b))
    if allsum is not None:
        s = allsum(s)
    bnorm = np.sqrt(s)
    b = b / bnorm
    return dict(rows=(S.indptr + 1).astype(np.int32), cols=(S.indices + 1).astype(np.int32), vals=np.ascontiguousarray(S.data), b=np.ascontiguousarray(b),
                goffset=(goff_nodes * 3).astype(np.int32), gn=gn, ndeg=3, D_own=D_own, bnorm=bnorm, cont=cont, n_ghost_nodes=int(ghosts.size))


def load_mesh(meshdir, nparts):
    """Global node coordinates and hex8 connectivity from the partition files (with -partdual ElmerGrid writes only partitioning.N; without
    halo elements every element is in exactly one part.k.elements, shared nodes are repeated in several part.k.nodes)."""
    pdir = os.path.join(meshdir, "partitioning.%d" % nparts)
    nd = np.concatenate([read_table(os.path.join(pdir, "part.%d.nodes" % (p + 1))) for p in range(nparts)])
    ids = nd[:, 0].astype(np.int64)
    nn = int(ids.max())
    xyz = np.zeros((nn, 3))
    xyz[ids - 1] = nd[:, 2:5]
    el = np.concatenate([read_table(os.path.join(pdir, "part.%d.elements" % (p + 1))) for p in range(nparts)])
    assert (el[:, 2] == 808).all()
    el = el[np.argsort(el[:, 0].astype(np.int64), kind="stable")]
    assert np.array_equal(el[:, 0].astype(np.int64), np.arange(1, el.shape[0] + 1)), "every element exactly once"
    return np.ascontiguousarray(xyz), np.ascontiguousarray(el[:, 3:11].astype(np.int64))


def dry(args):
    """CPU check of the per-rank construction: the owned rows of every rank, put together, equal the globally assembled, scaled system in
    continuous numbering."""
    from elmerfem_b200 import synth
    import scipy.sparse as sp
    d = tempfile.mkdtemp(prefix="c3dry_")
    dt = make_mesh(d, args.ex, args.ey, args.ez, args.nparts)
    md = os.path.join(d, "beam")
    xyz, elems = load_mesh(md, args.nparts)
    owner, files = ownership(md, args.nparts)
    print("mesh %d nodes %d elements, ElmerGrid %.1f s, owned nodes per partition %s" % (xyz.shape[0], elems.shape[0], dt, np.bincount(owner).tolist()))
    # pass 1: unscaled owned diagonals -> global D
    parts = [rank_system(md, args.nparts, r, owner, files, xyz, elems) for r in range(args.nparts)]
    Dg = np.concatenate([p["D_own"] for p in parts])
    bsum = 0.0
    for p in parts:
        bsum += float(np.dot(p["b"] * p["bnorm"], p["b"] * p["bnorm"]))
    parts = [rank_system(md, args.nparts, r, owner, files, xyz, elems, allgather_D=lambda Do, off: Dg, allsum=lambda s: bsum) for r in range(args.nparts)]
    gn = parts[0]["gn"]
    S = sp.vstack([sp.csr_matrix((p["vals"], p["cols"] - 1, p["rows"] - 1), shape=(p["rows"].size - 1, gn)) for p in parts]).tocsr()
    b = np.concatenate([p["b"] for p in parts])
    # global assembly in natural numbering, permuted
    rows, cols, diag = synth.crs_structure(xyz.shape[0], elems.astype(np.int32), 3)
    vals, rhs = synth.assemble(1, [1e9, 0.3, 0.0, -1e4, 0.0], xyz, elems.astype(np.int32), 3, rows, cols, uniform=False)
    A = synth.CRS(rows, cols, diag, vals, 3)
    fixed = np.flatnonzero(xyz[:, 0] <= 1e-12) + 1
    dofs = np.sort(np.concatenate([3 * (fixed - 1) + c + 1 for c in range(3)])).astype(np.int32)
    synth.dirichlet(A, rhs, dofs, 0.0, False)
    synth.scale_system(A, rhs)
    cont = parts[0]["cont"]
    perm = (np.repeat(cont * 3, 3) + np.tile(np.arange(3), cont.size))
    C = A.to_scipy().tocoo()
    G = sp.csr_matrix((C.data, (perm[C.row], perm[C.col])), shape=(gn, gn)); G.sort_indices()
    bg = np.zeros(gn); bg[perm] = rhs
    assert np.array_equal(G.indptr, S.indptr) and np.array_equal(G.indices, S.indices), "structure differs"
    err = np.abs(G.data - S.data).max() / np.abs(G.data).max()
    print("structure identical; values rel. diff %.2e; rhs rel. diff %.2e" % (err, np.abs(bg - b).max() / np.abs(bg).max()))
    assert err < 1e-12
    from oracle import halo_oracle as HO
    off = [int(v) for v in parts[0]["goffset"]]
    sends = [HO.send_lists_rank(p["rows"].astype(np.int64) - 1, p["cols"].astype(np.int64) - 1, off, r) for r, p in enumerate(parts)]
    for r in range(args.nparts):
        pl = HO.plan_rank(sends, off, r)
        print("rank %d: %d owned dofs, neighbours %s, sends %d, ghosts %d" % (r, off[r + 1] - off[r], pl["neigh"].tolist(), pl["send_idx"].size, pl["ghost_gid"].size))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ex", type=int, default=400); ap.add_argument("--ey", type=int, default=143); ap.add_argument("--ez", type=int, default=143)
    ap.add_argument("--nparts", type=int, default=0); ap.add_argument("--dry", action="store_true")
    ap.add_argument("--out", default=None); ap.add_argument("--steps", type=int, default=2); ap.add_argument("--precond", default="ilu0")
    ap.add_argument("--method", default="bicgstabl")
    args = ap.parse_args()
    if args.dry:
        return dry(args)
    import torch
    import torch.distributed as dist
    import elmerfem_b200 as B
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(minutes=40))
    shared = [None]
    t_grid = 0.0
    if rank == 0:
        d = tempfile.mkdtemp(prefix="c3_")
        t_grid = make_mesh(d, args.ex, args.ey, args.ez, world)
        shared = [os.path.join(d, "beam")]
    dist.broadcast_object_list(shared, src=0)
    md = shared[0]
    t0 = time.time()
    xyz, elems = load_mesh(md, world)
    owner, files = ownership(md, world)
    t_read = time.time() - t0

    def allsum(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda"); dist.all_reduce(t); return float(t.item())

    def allgather_D(D_own, off):
        n = int(off[-1])
        full = torch.zeros(n, dtype=torch.float64, device="cuda")
        full[int(off[rank]):int(off[rank + 1])] = torch.from_numpy(D_own).cuda()
        dist.all_reduce(full)
        return full.cpu().numpy()

    t0 = time.time()
    p = rank_system(md, world, rank, owner, files, xyz, elems, allgather_D=allgather_D, allsum=allsum)
    t_asm = time.time() - t0
    del xyz, elems
    n, nnz = p["rows"].size - 1, p["cols"].size
    assert nnz < 2 ** 31 and p["gn"] < 2 ** 31, "per-rank nnz / global dofs must fit the 32-bit CRS indices of Matrix_t"
    ids = [B.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    M = B.Matrix()
    M.comm_init(world, rank, ids[0])
    M.set_partition(p["gn"], p["rows"], p["cols"], p["goffset"], 1, 3)
    M.set_values(p["vals"])
    kw = dict(method=args.method, precond=args.precond, bicgstabl_l=4, tol=1e-8, maxit=3000)
    res = []
    for s in range(args.steps + 1):
        M.factorize() if args.precond != "diagonal" and args.precond != "none" else None
        torch.cuda.synchronize(); dist.barrier()
        r = M.solve(p["b"], **kw)
        res.append(r)
    st = res[-1]["stats"]
    dist.barrier()
    spmv_ms = M.time_matvec(20)
    dist.barrier()
    lu_ms = M.time_lu(10) if args.precond.startswith("ilu") else 0.0
    # true residual
    ax = M.matvec(res[-1]["x"])
    rr = allsum(float(np.sum((ax - p["b"]) ** 2))); bb = allsum(float(np.sum(p["b"] ** 2)))

    def red(v, op):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=op); return float(t.item())
    solve_ms = red(st["solve_ms"], dist.ReduceOp.MAX); factor_ms = red(st["factor_ms"], dist.ReduceOp.MAX)
    spmv_max = red(spmv_ms, dist.ReduceOp.MAX); lu_max = red(lu_ms, dist.ReduceOp.MAX)
    nnz_tot = red(nnz, dist.ReduceOp.SUM); nnz_max = red(nnz, dist.ReduceOp.MAX); n_max = red(n, dist.ReduceOp.MAX); n_min = red(n, dist.ReduceOp.MIN)
    sys.path.insert(0, ROOT)
    import bench
    hp = bench.halo_parity(M, dict(part=p), rank, world, dist)
    plan = M.halo_plan()
    neigh_max = red(plan["neigh"].size, dist.ReduceOp.MAX); ghost_max = red(plan["ghost_gid"].size, dist.ReduceOp.MAX)
    if rank == 0:
        peak = bench.peaks()[0]
        its = res[-1]["iters"]
        out = {"config": "C3: elasticity beam %dx%dx%d hex8, 3 dofs/node, %d dofs, %d nnz, ElmerGrid -partdual -metiskway %d, %s(l=4)+%s, tol 1e-8" %
                         (args.ex, args.ey, args.ez, p["gn"], int(nnz_tot), world, args.method, args.precond),
               "n_gpus": world, "rounds": its, "info": res[-1]["info"], "solve_ms": solve_ms, "rounds_per_s": its / (solve_ms / 1e3), "factor_ms": factor_ms,
               "spmv_ms": spmv_max, "spmv_gbs_per_gpu_max_rank": bench.spmv_bytes(n_max, nnz_max) / (spmv_max * 1e-3) / 1e9,
               "spmv_frac_of_hbm_peak": bench.spmv_bytes(n_max, nnz_max) / (spmv_max * 1e-3) / 1e9 / peak,
               "lu_ms": lu_max, "lu_frac_of_hbm_peak": (bench.lu_bytes(n_max, nnz_max) / (lu_max * 1e-3) / 1e9 / peak) if lu_max > 0 else None,
               "dofs_per_rank_min_max": [int(n_min), int(n_max)], "neighbours_max": int(neigh_max), "ghosts_max": int(ghost_max),
               "true_residual": float(np.sqrt(rr / bb)), "parity_halo": hp, "tri_mode": st["tri_mode"],
               "host_s": {"elmergrid_mesh_and_metis": t_grid, "read": t_read, "assemble_rank0": t_asm}}
        print(json.dumps(out))
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(out) + "\n")
    M.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- Krylov throughput of the B200 linear-solve path on BASELINE.json's workloads.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                    (the CPU restatement of Elmer's path, host cores)

Workload (config.workload):
  N = 1 : BASELINE.json configs[1]: steady heat equation, unit cube, 200^3 hex8 (8,120,601 dofs, 217,081,801 nnz),
          Elmer's default diagonal scaling, BiCGStab + ILU0, tol 1e-8, fp64.
  N > 1 : the same problem weak-scaled: 200 x 200 x 200*N elements, slab-partitioned along z (ElmerGrid
          `-partition 1 1 N`), one slab of ~8.1M dofs per GPU, block-Jacobi ILU0 as in Elmer's MPI path.

  --workload elasticity : BASELINE.json configs[4] (C5): linear elasticity, 3 dofs/node, 137 x 137 x (140 N - 1) hex8 beam cut into z-slabs,
          7,998,480 dofs per GPU, BiCGStab(l=4) + Jacobi; a step = the first --elas-rounds rounds (8 SpMV each), NOT a converged
          solve (config.workload says so); the dominant kernel is then the SpMV.  The reference arm always runs the heat workload.

A step is one IterSolver call on resident data: x = 0, BiCGStab+ILU0 to convergence.  `value` is BASELINE.json's figure, Krylov
iterations per second of the whole job (under weak scaling the global problem grows with N while an iteration should cost the same:
a flat `value` over N is perfect weak scaling); `mdof_iterations_per_s` (x global dofs) and the time to solution (`solve_s`,
`iterations_per_solve`: block-Jacobi ILU0 needs more iterations as partitions are added, exactly as Elmer's MPI path does) are
printed beside it.  Unless --no-c5 is given the same run also measures BASELINE configs[4] (C5: elasticity, BiCGStab(l=4)+Jacobi,
~8M dofs per GPU) and reports it as `c5_elasticity` inside the same JSON line.  At N > 1 every rank's halo index lists are
compared with the restatement of the reference's construction (`parity_halo`).  The ILU0 factorisation is done (and timed, `factor_ms`)
once per step outside the solve timer, for both arms.  `e2e` goes through b200_solve with pinned HOST
b/x (H2D + D2H inside the timed region).  The matrix (2.6 GB) is far larger than L2 (126 MB), so no
flush is needed between steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOL = 1e-8
MAXIT = 2000


# ------------------------------------------------------------------------------------------------
def spmv_bytes(n, nnz):
    """CRS-equivalent algorithmic bytes of one fp64 SpMV."""
    return 12 * nnz + 20 * n + 4


def lu_bytes(n, nnz):
    """one ILU0 application (forward + backward sweep)."""
    return 12 * nnz + 4 * (n + 1) + 4 * n + 24 * n


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def make_problem(args, rank, nranks, allreduce_sum=None):
    """Scaled system of this rank.  Returns dict(A=CRS local, b, gn, goffset, cols_global or None)."""
    from elmerfem_b200 import synth
    ne = args.ne
    if args.workload == "elasticity":
        # BASELINE configs[4] (C5): ~8M dofs per GPU, 3 dofs/node, beam along z cut into z-slabs
        nz = args.elas_layers
        part = synth.elasticity_slab(ne, ne, nz * nranks - 1, rank, nranks, allreduce_sum=allreduce_sum)
        name = "elasticity_beam_%dx%dx%d_hex8_3dof_z-slabs" % (ne, ne, nz * nranks - 1)
        if nranks > 1:
            return dict(A=None, b=part["b"], gn=part["gn"], part=part, name=name)
        n = part["rows"].size - 1
        rowid = np.repeat(np.arange(1, n + 1, dtype=np.int64), np.diff(part["rows"]))
        diag = (np.flatnonzero(part["cols"] == rowid) + 1).astype(np.int32)
        assert diag.size == n
        A = synth.CRS(part["rows"], part["cols"], diag, part["vals"], 3)
        return dict(A=A, b=part["b"], gn=n, part=None, name=name)
    if nranks == 1:
        A, b = synth.workload("heat", ne)
        return dict(A=A, b=b, gn=A.n, part=None, name="heat_cube_%d^3_hex8" % ne)
    part = synth.heat_slab(ne, ne, ne * nranks, rank, nranks, allreduce_sum=allreduce_sum)
    return dict(A=None, b=part["b"], gn=part["gn"], part=part, name="heat_slab_%dx%dx%d_hex8_z-slabs" % (ne, ne, ne * nranks))


def run_b200(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = measure(args, args.workload, rank, world, local, dist)
    if args.workload == "heat" and not args.no_c5:
        import copy
        a2 = copy.copy(args); a2.workload = "elasticity"; a2.ne = 137
        c5 = measure(a2, "elasticity", rank, world, local, dist)
        if rank == 0:
            keep = ["metric", "value", "unit", "mdof_iterations_per_s", "ms_per_step", "n_gpus", "steps", "scaling", "config", "roofline", "e2e", "gpu_launches",
                    "spmv_gbs", "spmv_frac_of_hbm_peak", "parity_halo", "allreduces_per_round", "launches_per_round"]
            out["c5_elasticity"] = {k: c5[k] for k in keep if k in c5}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def halo_parity(M, prob, rank, world, dist):
    """Integer parity inside the run (N > 1): this rank's halo lists, as the library built them through NCCL, against the numpy
    restatement of elmer_distribute_matrix (oracle/halo_oracle.py send_lists_rank / plan_rank; rocalution.cpp:64-372).  Checker only:
    runs after the timed regions."""
    import hashlib
    import torch
    from oracle import halo_oracle as HO
    p = prob["part"]
    rows0 = np.asarray(p["rows"], dtype=np.int64) - 1
    cols0 = np.asarray(p["cols"], dtype=np.int64) - 1
    off = [int(v) for v in p["goffset"]]
    mine = HO.send_lists_rank(rows0, cols0, off, rank)
    allsend = [None] * world
    dist.all_gather_object(allsend, mine)
    ref = HO.plan_rank(allsend, off, rank)
    got = M.halo_plan()
    ok = all(np.array_equal(np.asarray(got[k], dtype=np.int32), np.asarray(ref[k], dtype=np.int32)) for k in ["neigh", "send_ptr", "send_idx", "recv_ptr", "ghost_gid"])
    hh = hashlib.sha256(b"".join(np.ascontiguousarray(got[k], dtype=np.int32).tobytes() for k in ["neigh", "send_ptr", "send_idx", "recv_ptr", "ghost_gid"])).hexdigest()
    t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"result": "bit-exact" if t.item() == 1.0 else "MISMATCH", "ranks": world, "send_entries_rank0": int(got["send_idx"].size),
            "sha256_rank0": hh, "against": "oracle/halo_oracle.py (restatement of fem/src/rocalution.cpp:64-372), every rank"}


def measure(args, workload, rank, world, local, dist):
    import torch
    import elmerfem_b200 as B
    import ctypes as C

    def allsum(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    args = __import__("copy").copy(args); args.workload = workload
    prob = make_problem(args, rank, world, allsum if world > 1 else None)
    M = B.Matrix()
    t0 = time.time()
    if world == 1:
        A = prob["A"]
        M.set_structure(A.rows, A.cols, A.diag, 1, A.ndeg)
        n, nnz = A.n, A.nnz
        vals = A.vals
    else:
        ids = [B.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        M.comm_init(world, rank, ids[0])
        p = prob["part"]
        M.set_partition(p["gn"], p["rows"], p["cols"], p["goffset"], 1, p.get("ndeg", 1))
        n, nnz = p["rows"].size - 1, p["cols"].size
        vals = p["vals"]
    t_struct = time.time() - t0
    t0 = time.time()
    M.set_values(vals)
    t_upload = time.time() - t0
    gn = prob["gn"]
    b_host = torch.from_numpy(np.ascontiguousarray(prob["b"])).pin_memory()
    x_host = torch.zeros(n, dtype=torch.float64).pin_memory()
    nvec = M.vec_len()
    d_b = torch.zeros(nvec, dtype=torch.float64, device="cuda")
    d_b[:n].copy_(b_host)
    d_x = torch.zeros(nvec, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    elas = args.workload == "elasticity"
    if elas:   # C5: BiCGStab(l=4) + Jacobi; a step = the first `--elas-rounds` rounds (8 SpMV each), not a converged solve
        kw = dict(method="bicgstabl", precond="diagonal", bicgstabl_l=4, tol=TOL, maxit=args.elas_rounds)
    else:
        kw = dict(method="bicgstab", precond="ilu0", tol=TOL, maxit=MAXIT)
    label = "BiCGStab(l=4)+Jacobi, first %d rounds of 8 SpMV" % args.elas_rounds if elas else "BiCGStab+ILU0, tol 1e-8"

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if not elas:
            M.factorize()
        d_x.zero_()
        torch.cuda.synchronize()
        r = M.solve_device(d_b.data_ptr(), d_x.data_ptr(), **kw)
        return r

    def step_e2e():
        x_host.zero_()
        t = time.perf_counter()
        r = M.solve_host(b_host.data_ptr(), x_host.data_ptr(), **kw)      # b200_solve: H2D b, x; solve; D2H x
        return r, time.perf_counter() - t

    # ---- warm-up
    for _ in range(args.warmup):
        r = step_resident()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    # ---- timed: K steps, device time of the solve (CUDA events on the solve stream, inside the library)
    barrier()
    wall0 = time.perf_counter()
    solve_ms, factor_ms, iters, launches = 0.0, 0.0, 0, 0
    for _ in range(args.steps):
        r = step_resident()
        st = r["stats"]
        solve_ms += st["solve_ms"]; factor_ms += st["factor_ms"]; iters += r["iters"]; launches += st["launches"] + st["factor_launches"]
        assert r["info"] == 1 or (elas and r["info"] == 2), "solve did not converge: HUTI_INFO=%d" % r["info"]
    barrier()
    wall = time.perf_counter() - wall0
    # ---- e2e through the host-buffer entry point
    for _ in range(min(2, args.warmup)):
        step_e2e()
    barrier()
    e2e_s, e2e_iters, h2d, d2h = 0.0, 0, 0, 0
    for _ in range(args.steps):
        r2, dt = step_e2e()
        e2e_s += dt; e2e_iters += r2["iters"]; h2d = r2["stats"]["h2d"]; d2h = r2["stats"]["d2h"] + 0
        assert r2["info"] == 1 or (elas and r2["info"] == 2)
    barrier()
    # ---- one nonlinear iteration's worth of linear-solve work through the host-buffer ABI: upload Values, factorise, solve
    t_nl = time.perf_counter()
    M.set_values(vals)
    if not elas:
        M.factorize()
    x_host.zero_()
    r3 = M.solve_host(b_host.data_ptr(), x_host.data_ptr(), **kw)
    torch.cuda.synchronize()
    t_nl = time.perf_counter() - t_nl
    barrier()
    clk = clocks.stop() if rank == 0 else None
    # ---- kernel rooflines, measured live with CUDA events on the solve stream.  The barrier matters for N > 1: the halo
    # product waits for its neighbours, so a rank entering late (rank 0 has just stopped the clock sampler) would be
    # timed as SpMV time by the others (measured: 17.7 ms "per SpMV" at N = 8 without it).
    barrier()
    spmv_ms = M.time_matvec(20)
    barrier()
    lu_ms = M.time_lu(10) if not elas else 0.0
    tri_mode = M.stats()["tri_mode"]

    def maxr(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    maxr_nl = maxr(t_nl)
    solve_ms = maxr(solve_ms); factor_ms = maxr(factor_ms); e2e_s = maxr(e2e_s); spmv_ms_max = maxr(spmv_ms); lu_ms_max = maxr(lu_ms)
    launches = int(sumr(launches)); nnz_tot = sumr(float(nnz)); h2d = int(sumr(h2d)); d2h = int(sumr(d2h))
    # true residual of the answer: one more (halo-exchanging) SpMV through the host entry point, norms summed over ranks
    xs = d_x[:n].cpu().numpy()
    ax = M.matvec(xs)
    res_true = float(np.sqrt(sumr(float(np.sum((ax - prob["b"]) ** 2))) / sumr(float(np.sum(prob["b"] ** 2)))))

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        its = iters / (solve_ms / 1e3)
        e2e_its = e2e_iters / e2e_s
        bs, bl = spmv_bytes(n, nnz), lu_bytes(n, nnz)
        spmv_gbs = bs / (spmv_ms_max * 1e-3) / 1e9
        lu_gbs = bl / (lu_ms_max * 1e-3) / 1e9 if lu_ms_max > 0 else 0.0
        ipi = iters / args.steps
        # share of the solve spent in each kernel family (per iteration: 3 SpMV + 2 LU applications;
        # BiCGStab(l=4): 8 SpMV per round, no triangular solves with Jacobi)
        share_spmv = (8 if elas else 3) * spmv_ms_max * ipi / (solve_ms / args.steps)
        share_lu = 2 * lu_ms_max * ipi / (solve_ms / args.steps)
        dominant = "lu" if share_lu > share_spmv else "spmv"
        traffic = {}
        for tp in (os.path.join(ROOT, "profiles", "r02_traffic.json"), os.path.join(ROOT, "profiles", "r01_traffic.json")):
            if os.path.exists(tp):
                traffic = json.load(open(tp)).get(prob["name"], {})
                if traffic:
                    break
        lu_traffic = traffic.get({0: "lu_level", 3: "lu_wave", 4: "lu_lane"}.get(tri_mode, "lu"), traffic.get("lu"))
        roof = {"spmv": {"kernel": "k_spmv_sell", "bound": "hbm", "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak,
                         "traffic": traffic.get("spmv"), "bytes_per_launch": bs, "ms_per_launch": spmv_ms_max, "share_of_step": share_spmv, "peak_source": peak_src},
                "lu": {"kernel": {0: "k_sptrsv (level sweeps, L then U)", 3: "k_wave (wave tiles, L then U, + layout conversion)",
                                  4: "k_lane (lane tiles, L then U, + layout conversion)"}.get(tri_mode, "k_sptrsv"), "bound": "hbm", "achieved": lu_gbs, "peak": peak, "unit": "GB/s", "frac": lu_gbs / peak,
                       "traffic": lu_traffic, "bytes_per_launch": bl, "ms_per_launch": lu_ms_max, "share_of_step": share_lu, "peak_source": peak_src}}
        out = {
            "metric": ("fp64 Krylov iterations/s (BiCGStab(l=4)+Jacobi rounds of 8 SpMV, elasticity ~8M dof per GPU); SpMV GB/s and fraction of the HBM roofline beside it" if elas else
                       "fp64 Krylov iterations/s (BiCGStab+ILU0, heat 200^3 per GPU); SpMV GB/s and fraction of the HBM roofline beside it"),
            "value": its, "unit": "rounds/s" if elas else "iterations/s",
            "iters_per_s": its, "mdof_iterations_per_s": its * gn / 1e6, "spmv_gbs": spmv_gbs * world, "spmv_frac_of_hbm_peak": spmv_gbs / peak,
            "solve_s": solve_ms / args.steps / 1e3, "iterations_per_solve": ipi,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": solve_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": prob["name"] + ", " + label + ", Linear System Scaling on", "global_dofs": int(gn),
                       "global_nnz": int(nnz_tot), "dofs_per_gpu": int(n), "iterations_per_solve": ipi,
                       "time_to_solution": ("solve_s = ms_per_step: one converged solve of the GLOBAL problem; under weak scaling it grows with iterations_per_solve "
                                            "(block-Jacobi ILU0 per partition, as in Elmer's MPI path, and a longer domain), while `value` (per-iteration rate) should stay flat"),
                       "l2": "inputs (%.1f GB matrix per GPU) larger than L2; no flush" % (12.0 * nnz / 1e9), "parallelism": "row partition, z-slabs x%d" % world},
            "roofline": roof[dominant], "roofline_spmv": roof["spmv"], "roofline_lu": roof["lu"],
            "e2e": {"value": e2e_its, "unit": "rounds/s" if elas else "iterations/s", "mdof_iterations_per_s": e2e_its * gn / 1e6, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": launches, "factor_ms": factor_ms / args.steps, "upload_values_s": t_upload, "structure_s": t_struct,
            "nonlinear_iteration_ms": {"value": maxr_nl * 1e3, "what": "b200_set_values (host Values -> device, %s) + factorisation + b200_solve with host b/x, wall clock, max over ranks" %
                                       ("page-locked once, B200_PIN_VALUES=1" if os.environ.get("B200_PIN_VALUES") == "1" else "pageable"), "iterations": r3["iters"]},
            "iters_per_s_incl_factor": iters / ((solve_ms + factor_ms) / 1e3), "wall_s_timed_region": wall,
            "true_residual": res_true, "clocks": clk,
        }
        if elas:
            # ndeg = 3: the SpMV reads one column index per group of 3 entries (as the reference's ndeg loop does), so it moves
            # fewer bytes than the CRS-equivalent figure `achieved` is defined on: say so, with the real figure beside it
            real = (8.0 + 4.0 / 3.0) * nnz + 20.0 * n + 4
            for key in ("spmv",):
                roof[key]["real_bytes_estimate"] = real
                roof[key]["achieved_real"] = real / (spmv_ms_max * 1e-3) / 1e9
                roof[key]["frac_real"] = roof[key]["achieved_real"] / peak
                roof[key]["note"] = ("achieved/frac use the CRS-equivalent algorithmic bytes (12 nnz + 20 n); the block-column kernel reads one index per 3 entries, "
                                     "so the bytes it really moves are ~(8 + 4/3) nnz + 20 n: achieved_real / frac_real")
            out.pop("roofline_lu"); out["roofline"] = roof["spmv"]; out.pop("factor_ms"); out.pop("iters_per_s_incl_factor")
            out["true_residual_note"] = "not converged by design (fixed number of rounds)"
        gp = os.path.join(ROOT, "tests", "golden", "c2_device_order.json")
        if world == 1 and not elas and args.ne == 200 and os.path.exists(gp):
            gold = json.load(open(gp))
            out["config"]["iterations_reference"] = {
                "what": "iterations of the CPU restatement of Elmer's BiCGStab+ILU0 on this very system under four summation orders of ddot (tests/golden/c2_device_order.json); "
                        "the device solve is bit-identical to `device order` (tests/test_gpu_bitwise.py)",
                **{gold[k]["what"]: gold[k]["iters"] for k in ("order_0", "order_1", "order_2", "order_3") if k in gold}}
            out["config"]["iterations_device"] = ipi
        if elas:
            out["allreduces_per_round"] = 2 * 4 + 2 if world > 1 else 0
            out["launches_per_round"] = launches / max(1, iters) / world
        if world == 1 and not args.no_cpu_baseline and not elas:
            out["cpu_baseline"] = cpu_baseline(prob, budget_s=args.cpu_budget)
    if world > 1:
        hp = halo_parity(M, prob, rank, world, dist)
        if rank == 0:
            out["parity_halo"] = hp
    M.close()
    del d_b, d_x
    torch.cuda.empty_cache()
    if dist is not None:
        dist.barrier()
    return out


# ------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample(O, A, b, ilu, maxit):
    t = time.perf_counter()
    r = O.itersolve(A, b, method="bicgstab", precond="ilu0", ilu=ilu, tol=TOL, maxit=maxit)
    return r, time.perf_counter() - t


def cpu_baseline(prob, budget_s=20.0):
    """The oracle (literal C++ restatement of Elmer's CPU path: OpenMP SpMV/vector loops, serial ILU0
    solves exactly as the reference) on this box's host cores, on a bounded sample of the same system:
    the first m BiCGStab+ILU0 iterations."""
    from oracle import oracle as O
    A, b = prob["A"], prob["b"]
    cores = host_cores()
    O.set_threads(cores)
    t = time.perf_counter(); ilu = O.ilu0(A); t_f = time.perf_counter() - t
    r, dt = cpu_sample(O, A, b, ilu, 2)
    m = int(max(2, min(MAXIT, budget_s / (dt / 2))))
    r, dt = cpu_sample(O, A, b, ilu, m)
    its = r["iters"] if r["info"] == 1 else min(r["iters"], m)
    return {"value": its / dt, "unit": "iterations/s", "mdof_iterations_per_s": its / dt * A.n / 1e6, "cores": cores, "kind": "port",
            "sample": "first %d BiCGStab+ILU0 iterations of the same %d-dof system (%.1f s); ILU0 factor %.2f s not included" % (its, A.n, dt, t_f),
            "factor_s": t_f}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    prob = make_problem(args, 0, 1)
    A, b = prob["A"], prob["b"]
    cores = host_cores()            # torchrun exports OMP_NUM_THREADS=1: ask for every core of the box explicitly
    O.set_threads(cores)
    t = time.perf_counter(); ilu = O.ilu0(A); t_f = time.perf_counter() - t
    r, dt = cpu_sample(O, A, b, ilu, 2)
    per_it = dt / 2
    total = args.steps + args.warmup
    m = int(max(2, min(MAXIT, args.ref_budget / total / per_it)))
    for _ in range(args.warmup):
        cpu_sample(O, A, b, ilu, m)
    T, its = 0.0, 0
    for _ in range(args.steps):
        r, dt = cpu_sample(O, A, b, ilu, m)
        T += dt; its += min(r["iters"], m)
    v = its / T
    sample = "each step = first %d BiCGStab+ILU0 iterations of the same %d-dof system; ILU0 factor (%.2f s) outside the timer as in the GPU arm" % (m, A.n, t_f)
    out = {"impl": "reference", "metric": "fp64 Krylov iterations/s (BiCGStab+ILU0, heat 200^3 per GPU); SpMV GB/s and fraction of the HBM roofline beside it",
           "value": v, "unit": "iterations/s", "iters_per_s": v, "mdof_iterations_per_s": v * A.n / 1e6, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": T / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": prob["name"] + ", BiCGStab+ILU0, tol 1e-8, Linear System Scaling on", "global_dofs": int(A.n), "global_nnz": int(A.nnz)},
           "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample, "factor_s": t_f},
           "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": ("reference = C++ restatement of Elmer's CPU Krylov path (no Fortran compiler in this image), all %d host cores. It always solves the ONE-partition "
                    "problem (8.1 M dofs): at --gpus N > 1 the GPU arm iterates on an N times larger global system, so value / reference_value then UNDERSTATES the "
                    "speed-up per unit of work and is not a like-for-like ratio; only the N = 1 ratio is") % cores,
           "same_problem_as_gpu_arm": args.gpus == 1}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ne", type=int, default=None, help="elements per edge of the cross-section (default 200 = BASELINE configs[1]; elasticity: 137)")
    ap.add_argument("--workload", default="heat", choices=["heat", "elasticity"], help="heat = BASELINE configs[1] (default); elasticity = configs[4] (C5 weak scaling)")
    ap.add_argument("--elas-layers", type=int, default=140, help="elasticity: node layers per GPU along the beam")
    ap.add_argument("--elas-rounds", type=int, default=40, help="elasticity: BiCGStab(4) rounds per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="heat workload only: skip the C5 elasticity measurement that is otherwise reported as c5_elasticity")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--ref-budget", type=float, default=150.0)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.ne is None:
        args.ne = 137 if args.workload == "elasticity" else 200
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

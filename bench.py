#!/usr/bin/env python
"""Headline benchmark: curl-curl SpMV on the Dey-Mittra pillbox (BASELINE.json configs[1]).

One "step" = one apply y = curlCurl * x of the assembled operator over the whole grid
(MxCrsMatrix::apply, reference src/MxCrsMatrix.cpp:347-353) on a block of --nvec vectors.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 256] [--nvec 1] [--layout dict|sell]
  python bench.py --impl reference ...      # the reference's CPU path (oracle port) on host cores

Multi-GPU: one process per GPU (torchrun); the grid is x-slab decomposed (strong scaling, the
operator is fixed), ghost planes move with NCCL send/recv inside mxg_crs_apply.
Prints ONE JSON line on rank 0.
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


def crs_bytes(nnz, nrows, b, is_complex):
    """CRS-equivalent algorithmic bytes of one apply."""
    return (20 * nnz + nrows * (4 + 32 * b)) if is_complex else (12 * nnz + nrows * (4 + 16 * b))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.time(), line.strip()))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [s for s in self.samples if t0 - 0.15 <= s[0] <= t1 + 0.15] or self.samples[-3:]
        for _, line in rows:
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                for nm, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_operator(size, workload):
    """Generate the global operator with the (test-infrastructure) generator. bench.py is allowed to
    use oracle/ for input generation and the cpu_baseline leg only; the timed GPU path never touches it."""
    from oracle import oracle as orc
    t = time.time()
    if workload == "pillbox":
        sim = orc.pillbox(size)
    elif workload == "vacuum":
        sim = orc.vacuum(size)
    else:
        raise SystemExit("unknown workload " + workload)
    op = sim.op("curlCurl")
    rowptr, col, val = op.arrays()
    rg, cg = op.maps()
    info = {"gen_s": round(time.time() - t, 2), "n_global": int(sim.num_global("bfield"))}
    return op, rowptr, col, val, rg, info


def slab_ranges(row_gids, n_global, nranks, size):
    """x-slab partition: every rank owns whole x-planes of the (N+1)^3 node grid, i.e. one contiguous GID
    range (GID = comp + 3*((x*(Ny+1)+y)*(Nz+1)+z), MxGrid.h:96-114). Cut planes are chosen so the DOF counts
    are balanced (the PEC mask makes equal-width slabs very uneven)."""
    plane = n_global // (size + 1)
    first_of_plane = np.searchsorted(row_gids, np.arange(size + 2) * plane)
    cuts = [0]
    for r in range(1, nranks):
        target = r * len(row_gids) / nranks
        x = int(np.argmin(np.abs(first_of_plane - target)))
        cuts.append(int(first_of_plane[x]))
    cuts.append(len(row_gids))
    return cuts


def run_reference(args):
    from oracle import oracle as orc
    op, rowptr, col, val, rg, info = build_operator(args.size, args.workload)
    b = args.nvec
    rng = np.random.default_rng(12345)
    x = np.asfortranarray(rng.uniform(-1, 1, size=(op.ncols, b)))
    threads = orc.lib().mxo_num_threads()
    for _ in range(max(args.warmup, 1)):
        op.apply(x, threads)
    t0 = time.time()
    for _ in range(args.steps):
        op.apply(x, threads)
    dt = (time.time() - t0) / args.steps
    B = crs_bytes(op.nnz, op.nrows, b, op.is_complex)
    gbs = B / dt / 1e9
    print(json.dumps({
        "impl": "reference", "metric": "curlcurl_spmv_crs_equiv_gbs", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s-%d curlCurl SpMV (Dey-Mittra, %d rows, %d nnz)" % (args.workload, args.size, op.nrows, op.nnz),
                   "nvec": b, "layout": "crs (Epetra order)", "partition": "rows over %d host threads" % threads},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": "%d full-operator applies (Epetra-order CSR, OpenMP static rows)" % args.steps},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def upload_matrix(mx, mat, rmap, cmap):
    rowptr, col, val = mat.arrays()
    _, cg = mat.maps()
    return mx.MxCrsMatrix.from_csr(rmap, cmap, rowptr, cg[col], val)


def run_eigensolve(mx, ctx, args):
    """Lowest `nev` eigenpairs of vecLapl b = k^2 mRhs b (MxMagWaveOp.cpp:183-241) with the GPU multigrid
    V-cycle as preconditioner. The level operators are re-discretisations on halved grids
    (MxEMSimHierarchy.cpp:37-76) generated on the host; everything timed runs on the GPU."""
    from oracle import oracle as orc
    make = orc.pillbox if args.workload == "pillbox" else orc.vacuum
    sizes = [args.size]
    while sizes[-1] % 2 == 0 and sizes[-1] // 2 >= 8 and len(sizes) < 6:
        sizes.append(sizes[-1] // 2)
    t = time.time()
    sims = [make(n) for n in sizes]
    maps, ops = [], []
    for s in sims:
        m = s.op("vecLapl")
        rg, _ = m.maps()
        maps.append(mx.MxMap(ctx, s.num_global("bfield"), rg))
        ops.append(upload_matrix(mx, m, maps[-1], maps[-1]))
        del m
    R, P = [], []
    for l in range(len(sims) - 1):
        p = orc.interpolator(sims[l + 1], sims[l])
        P.append(upload_matrix(mx, p, maps[l], maps[l + 1]))
        R.append(upload_matrix(mx, p.transpose(scale=0.125), maps[l + 1], maps[l]))
        del p
    fa = sims[0].fracs("bfield")
    md = mx.MxMultiVector(maps[0], 1)
    md.from_host(fa)
    divB = sims[0].op("divB")
    pmap = mx.MxMap(ctx, sims[0].num_global("psifield"), divB.maps()[0])
    D = upload_matrix(mx, divB, pmap, maps[0])
    setup_s = time.time() - t
    del sims
    prec = mx.MxGeoMultigridPrec(ctx, ops, R, P, smoother_sweeps=2, cycles=1)
    solver = mx.MxSolver(ctx, ops[0], m_diag=md, prec=prec, nev=args.nev, block_size=args.block, tol=args.tol, max_iters=200)
    l0 = ctx.launch_count()
    ev = solver.solve()
    res, div = solver.check(D)
    nev = args.nev
    maxwell = [float(e) for e, d in zip(ev, div[:nev]) if d < 1e-5]
    return {"metric": "seconds_to_%d_eigenpairs" % nev, "value": solver.seconds, "unit": "s", "iterations": int(solver.iterations),
            "converged": int(solver.converged), "block": int(args.block), "tol": args.tol, "levels": sizes,
            "operator_applies": int(solver.apply_a), "vcycles": int(solver.apply_prec), "gpu_launches": int(ctx.launch_count() - l0),
            "eigenvalues": [float(e) for e in ev], "max_rel_residual": float(solver.residuals[:nev].max()),
            "reference_residual_check": float(res[:nev].max()), "divergence_free_modes": maxwell,
            "host_setup_s": round(setup_s, 1),
            "note": "pencil (vecLapl, dmA); modes with |div(M b)|/|M b| < 1e-5 are the Maxwell modes, the rest are "
                    "the grad-div modes the reference removes by projection (MxMagWaveOp.cpp:893-924)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="mxgpu")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--workload", default="pillbox")
    ap.add_argument("--nvec", type=int, default=1)
    ap.add_argument("--no-sweep", action="store_true", help="skip the informational 4- and 10-vector block applies")
    ap.add_argument("--layout", default="dict", choices=["dict", "sell"])
    ap.add_argument("--order", choices=["ref", "soa"], default="ref",
                    help="device ordering of the vectors: reference (ascending GID) or component-major (experimental, 1 GPU)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the eigensolve leg (seconds to nev eigenpairs)")
    ap.add_argument("--nev", type=int, default=10)
    ap.add_argument("--block", type=int, default=16)
    ap.add_argument("--tol", type=float, default=1e-8)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return

    import maxwell_b200 as mx
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)  # control plane only

    ctx = mx.Context(local_rank)
    if world > 1:
        ids = [mx.Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(rank, world, ids[0])

    # ---- operator: generated once (rank 0) and shared through /dev/shm --------------------
    shm = "/dev/shm/mxbench_%s" % os.environ.get("MASTER_PORT", str(os.getpid()))
    if rank == 0:
        op, rowptr, col, val, rg, info = build_operator(args.size, args.workload)
        if world > 1:
            os.makedirs(shm, exist_ok=True)
            for name, arr in (("rowptr", rowptr), ("col", col), ("val", val), ("rg", rg)):
                np.save(os.path.join(shm, name + ".npy"), arr)
            with open(os.path.join(shm, "info.json"), "w") as f:
                json.dump(info, f)
    if world > 1:
        dist.barrier()
        if rank != 0:
            rowptr = np.load(os.path.join(shm, "rowptr.npy"), mmap_mode="r")
            col = np.load(os.path.join(shm, "col.npy"), mmap_mode="r")
            val = np.load(os.path.join(shm, "val.npy"), mmap_mode="r")
            rg = np.load(os.path.join(shm, "rg.npy"), mmap_mode="r")
            with open(os.path.join(shm, "info.json")) as f:
                info = json.load(f)
            op = None
    n_global = info["n_global"]
    nrows_g, nnz_g = len(rg), int(rowptr[-1])
    is_complex = np.iscomplexobj(val)
    b = args.nvec

    cuts = slab_ranges(np.asarray(rg), n_global, world, args.size)
    r0, r1 = cuts[rank], cuts[rank + 1]
    my_gids = np.ascontiguousarray(rg[r0:r1])
    p0, p1 = int(rowptr[r0]), int(rowptr[r1])
    my_rowptr = np.ascontiguousarray(rowptr[r0:r1 + 1]) - p0
    my_cols = np.asarray(rg)[np.asarray(col[p0:p1])]  # local col index of the global op -> GID
    my_vals = np.ascontiguousarray(val[p0:p1])

    t = time.time()
    if args.order == "soa" and world > 1:
        raise SystemExit("--order soa (component-major device ordering) is single-rank only")
    bmap = mx.MxMap(ctx, n_global, my_gids, components=3 if args.order == "soa" else 1)
    layout = mx.LAYOUT_DICT if args.layout == "dict" else mx.LAYOUT_SELL
    A = mx.MxCrsMatrix.from_csr(bmap, bmap, my_rowptr, my_cols, my_vals, layout=layout)
    stats = A.stats()
    build_s = time.time() - t
    del my_cols, my_vals

    x = mx.MxMultiVector(bmap, b, is_complex)
    y = mx.MxMultiVector(bmap, b, is_complex)
    x.random(12345)

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing ---------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        A.apply(x, y)
    barrier()
    l0 = ctx.launch_count()
    w0 = time.time()
    ctx.event_record(0)
    for _ in range(args.steps):
        A.apply(x, y)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    barrier()
    w1 = time.time()
    launches = ctx.launch_count() - l0
    if world > 1:
        import torch
        tt = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt[0])
    ms_step = ms / args.steps

    # per-kernel split (1 GPU): averaged over a few profiled applies
    # (several ranks: dict_ms = interior rows incl. the final wait, sell_ms = boundary rows, pre_ms = NCCL exchange)
    split = None
    acc = {"dict_ms": 0.0, "sell_ms": 0.0, "pre_ms": 0.0, "total_ms": 0.0}
    reps = 20
    for _ in range(reps):
        r = A.apply_timed(x, y)
        for k in acc:
            acc[k] += r[k] / reps
    split = acc
    if world > 1:
        split = {"interior_ms": acc["dict_ms"], "boundary_ms": acc["sell_ms"], "nccl_ms": acc["pre_ms"], "total_ms": acc["total_ms"]}

    # ---- block applies (b in {1, 4, 10}); informational, same operator and generator ----
    sweep = None
    if world == 1 and b == 1 and not args.no_sweep:
        sweep = {}
        for nb in (4, 10):
            xb = mx.MxMultiVector(bmap, nb, is_complex)
            yb = mx.MxMultiVector(bmap, nb, is_complex)
            xb.random(12345)
            for _ in range(3):
                A.apply(xb, yb)
            ctx.sync()
            ctx.event_record(0)
            for _ in range(30):
                A.apply(xb, yb)
            ctx.event_record(1)
            t_nb = ctx.event_elapsed_ms(0, 1) / 30
            sweep[str(nb)] = {"ms_per_apply": t_nb, "ms_per_vector": t_nb / nb,
                              "crs_equiv_gbs": crs_bytes(nnz_g, nrows_g, nb, is_complex) / (t_nb * 1e-3) / 1e9}
            del xb, yb

    # ---- e2e: host buffers through the C ABI (H2D of x, apply, D2H of y every step) ------------
    n_loc = r1 - r0
    dt = np.complex128 if is_complex else np.float64
    xh = mx.pinned_array((n_loc, b), dt)
    yh = mx.pinned_array((n_loc, b), dt)
    xh[...] = x.to_host()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        x.from_host(xh); A.apply(x, y); y.to_host(yh)
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        x.from_host(xh)
        A.apply(x, y)
        y.to_host(yh)
    barrier()
    e2e_serial_s = (time.time() - t0) / e2e_steps
    # the same steps through the host-buffer entry point, which pipelines upload / apply / download of successive
    # items over two device slots (every item still crosses the bus both ways inside the timed region)
    e2e_s, e2e_mode = e2e_serial_s, "serial upload -> apply -> download per step"
    if b == 1 and args.order == "ref":
        yhs = [yh, mx.pinned_array((n_loc, b), dt)]
        x1 = xh[:, 0]
        A.apply_host_batch([x1, x1], [yhs[0][:, 0], yhs[1][:, 0]])
        same_e2e = bool(np.array_equal(yhs[0], yhs[1]) and np.array_equal(yhs[0], y.to_host()))
        barrier()
        t0 = time.time()
        A.apply_host_batch([x1] * e2e_steps, [yhs[i & 1][:, 0] for i in range(e2e_steps)])
        barrier()
        e2e_pipe_s = (time.time() - t0) / e2e_steps
        if same_e2e and e2e_pipe_s < e2e_serial_s:
            e2e_s, e2e_mode = e2e_pipe_s, "mxg_crs_apply_host_batch (pipelined over two device slots)"
    if world > 1:
        tt = torch.tensor([e2e_s, e2e_serial_s], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s, e2e_serial_s = float(tt[0]), float(tt[1])

    if rank == 0:
        sampler.stop()

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        threads = orc.lib().mxo_num_threads()
        xc = x.to_host()
        op.apply(xc, threads)
        t0 = time.time()
        n_cpu = 0
        while n_cpu < 3 or (time.time() - t0 < args.cpu_seconds and n_cpu < 1000):
            yc = op.apply(xc, threads)
            n_cpu += 1
        cpu_dt = (time.time() - t0) / n_cpu
        # parity spot check on the full-size operator while we are here: bit-exact
        A.apply(x, y)
        same = bool(np.array_equal(yc, y.to_host()))
        cpu = {"value": crs_bytes(nnz_g, nrows_g, b, is_complex) / cpu_dt / 1e9, "unit": "GB/s", "cores": threads,
               "kind": "port", "sample": "%d full-operator applies, %.1f s" % (n_cpu, cpu_dt * n_cpu),
               "ms_per_apply": cpu_dt * 1e3, "gpu_bit_exact_vs_cpu": same}

    # ---- eigensolve leg (BASELINE metric part 2): seconds to `nev` eigenpairs, multigrid-preconditioned ----
    solve = None
    if rank == 0 and world == 1 and not args.no_solve and not is_complex:
        del A, x, y
        solve = run_eigensolve(mx, ctx, args)

    if rank == 0:
        B = crs_bytes(nnz_g, nrows_g, b, is_complex)
        peak, peak_src = measured_peaks()
        gbs = B / (ms_step * 1e-3) / 1e9
        esz = 16 if is_complex else 8
        layout_bytes = stats["device_bytes"] + stats["rows"] * esz * b * 2  # this rank's streamed bytes per apply
        roof = {"bound": "hbm", "achieved": gbs / world, "peak": peak, "unit": "GB/s", "frac": gbs / world / peak,
                "traffic": None, "peak_source": peak_src,
                "note": "achieved = CRS-equivalent bytes (12*nnz + rows*(4+16b)) / time per GPU; the dictionary layout "
                        "streams far fewer bytes, see layout_gbs and profiles/",
                "layout_bytes_per_apply": int(layout_bytes),
                "layout_gbs": layout_bytes / (ms_step * 1e-3) / 1e9}
        if split:
            roof["kernel_ms"] = split
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of this workload
        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                tr = json.load(f).get("%s-%d-%s-%d" % (args.workload, args.size, args.layout, b))
            if tr and world == 1:
                roof["traffic"] = tr["dram_bytes"]
                # what actually crosses the HBM interface: far below the roofline, the compressed layout leaves the
                # kernel bound by L1 data-pipe wavefronts (profiles/README_r01.md)
                roof["dram_gbs"] = tr["dram_bytes"] / (ms_step * 1e-3) / 1e9
                roof["dram_frac"] = roof["dram_gbs"] / peak
                roof["traffic_source"] = ("profiles/r01_traffic.json (%s; ncu capture of the plain row assignment -- the "
                                          "interleaved kernel streams the same arrays)" % tr["kernel"])
        except Exception:
            pass
        out = {
            "metric": "curlcurl_spmv_crs_equiv_gbs", "value": gbs, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c128" if is_complex else "f64", "data": "synthetic",
            "config": {"workload": "%s-%d curlCurl SpMV (Dey-Mittra, %d rows, %d nnz)" % (args.workload, args.size, nrows_g, nnz_g),
                       "nvec": b, "layout": args.layout, "l2": "inputs_larger_than_l2" if layout_bytes > 126e6 else "fits_l2",
                       "partition": "x-slab x%d" % world, "gen_s": info["gen_s"], "layout_build_s": round(build_s, 2),
                       "vector_order": "component-major" if args.order == "soa" else "reference (ascending GID)"},
            "layout": stats,
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": B / e2e_s / 1e9, "unit": "GB/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": int(nrows_g * esz * b), "d2h_bytes_per_step": int(nrows_g * esz * b),
                    "mode": e2e_mode, "serial_ms_per_step": e2e_serial_s * 1e3},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(w0, w1),
            "wall_s_timed": w1 - w0,
            "eigensolve": solve,
            "block_applies": sweep,
        }
        print(json.dumps(out))
        if world > 1:
            import shutil
            shutil.rmtree(shm, ignore_errors=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

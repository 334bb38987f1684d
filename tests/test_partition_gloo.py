"""World-size-2 (and 3) CPU tests of the slab partition + halo plan: each rank owns an x-slab of rows,
exchanges ghost values with point-to-point messages (gloo here, NCCL send/recv on the GPU) and applies
its block; the stitched result must equal the global apply bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maxwell_b200.partition import HaloPlan, local_block, slab_cuts


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, workload, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    sim = orc.pillbox(n) if workload == "pillbox" else orc.vacuum(n)
    op = sim.op("curlCurl")
    rowptr, col, val = op.arrays()
    rg, _ = op.maps()
    n_global = sim.num_global("bfield")
    cuts = slab_cuts(rg, n_global, n + 1, world)
    r0, r1 = cuts[rank], cuts[rank + 1]
    my_gids = rg[r0:r1]
    lrp, lcol, lval = local_block(rowptr, rg[col], val, r0, r1)
    ranges = [(int(rg[cuts[q]]), int(rg[cuts[q + 1] - 1])) if cuts[q + 1] > cuts[q] else (0, -1) for q in range(world)]
    plan = HaloPlan(my_gids, lcol, ranges)

    # every rank publishes the GIDs it needs; owners answer (same protocol as mxg_spmv.cu: planHalo)
    need = [None] * world
    dist.all_gather_object(need, plan.ghosts)
    rng = np.random.default_rng(99)
    x_global = rng.uniform(-1, 1, len(rg))                  # identical on every rank (global-id keyed)
    x_local = x_global[r0:r1].copy()
    ghost_vals = np.zeros(len(plan.ghosts))
    reqs = []
    for q in range(world):
        if q == rank:
            continue
        mine = need[q][(need[q] >= ranges[rank][0]) & (need[q] <= ranges[rank][1])]
        if len(mine):
            buf = torch.from_numpy(x_local[plan.send_indices(my_gids, mine)].copy())
            reqs.append(dist.isend(buf, dst=q))
    recv_bufs = {}
    for q, (start, count) in plan.recv.items():
        recv_bufs[q] = torch.empty(count, dtype=torch.float64)
        reqs.append(dist.irecv(recv_bufs[q], src=q))
    for r in reqs:
        r.wait()
    for q, (start, count) in plan.recv.items():
        ghost_vals[start:start + count] = recv_bufs[q].numpy()
    xe = plan.extended_x(x_local, ghost_vals)
    y_local = orc.csr_apply(lrp, (plan.ext_col + plan.g_lo).astype(np.int32), lval, xe)
    y_ref = op.apply(x_global)[r0:r1]
    ok = np.array_equal(y_local, y_ref)
    # reductions: local partial + all_reduce == global dot (K5/K6 of SURVEY 2.3)
    part = torch.tensor([float(np.dot(x_local, y_local))], dtype=torch.float64)
    dist.all_reduce(part)
    ok_dot = abs(part.item() - float(np.dot(x_global, op.apply(x_global)))) <= 1e-9 * abs(part.item())
    np.save(os.path.join(out_dir, "r%d.npy" % rank),
            np.array([ok, ok_dot, len(plan.ghosts), plan.g_lo, plan.g_hi, r1 - r0, len(plan.recv)], dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("workload,n,world", [("vacuum", 8, 2), ("pillbox", 16, 2), ("pillbox", 12, 3)])
def test_slab_halo_exchange_matches_global_apply(tmp_path, workload, n, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, workload, n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok, ok_dot, nghost, glo, ghi, nloc, npeers = np.load(tmp_path / ("r%d.npy" % r))
        assert ok == 1, "rank %d: slab apply differs from the global apply" % r
        assert ok_dot == 1
        assert nloc > 0 and nghost > 0
        if workload == "vacuum":
            # periodic wrap: both neighbours exist even with 2 ranks; ghosts on both sides
            assert glo + ghi == nghost and npeers == world - 1


def test_slab_cuts_are_plane_aligned_and_balanced():
    from oracle import oracle as orc
    sim = orc.pillbox(24)
    rg = sim.map("bfield")
    n_global = sim.num_global("bfield")
    plane = n_global // 25
    for P in (2, 4, 8):
        cuts = slab_cuts(rg, n_global, 25, P)
        assert cuts[0] == 0 and cuts[-1] == len(rg) and all(b >= a for a, b in zip(cuts, cuts[1:]))
        for c in cuts[1:-1]:
            assert rg[c] // plane != rg[c - 1] // plane        # cut falls between two x-planes
        sizes = np.diff(cuts)
        assert sizes.max() <= 1.6 * len(rg) / P


def test_halo_plan_rejects_unowned_columns():
    my = np.array([10, 11, 12], dtype=np.int64)
    with pytest.raises(ValueError):
        HaloPlan(my, np.array([10, 50], dtype=np.int64), [(10, 12), (20, 30)])
    plan = HaloPlan(my, np.array([10, 25, 5, 12], dtype=np.int64), [(10, 12), (0, 9), (20, 30)])
    assert plan.g_lo == 1 and plan.g_hi == 1
    assert list(plan.ext_col) == [0, 3, -1, 2]
    assert plan.recv == {1: (0, 1), 2: (1, 1)}

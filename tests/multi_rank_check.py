"""Multi-GPU parity script, launched with torchrun by tests/test_gpu_multi.py (one process per GPU).
Checks the NCCL halo exchange inside mxg_crs_apply and the all-reduced block reductions against the
oracle's global (single-process) results. P-invariance: x is generated from global ids, so every rank
count sees the same vector."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist  # noqa: E402

import maxwell_b200 as mx  # noqa: E402
from maxwell_b200.partition import local_block, slab_cuts  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = mx.Context(local_rank)
    ids = [mx.Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(rank, world, ids[0])
    assert ctx.myPID() == rank and ctx.numProc() == world

    cases = [("vacuum", orc.vacuum(16), "curlCurl", 1), ("pillbox", orc.pillbox(24), "curlCurl", 3),
             ("pillbox-vecLapl", orc.pillbox(20), "vecLapl", 2),
             ("bloch", orc.vacuum(12, phase_shifts=(0.4, -1.3, 2.2)), "curlCurl", 2)]
    for label, sim, name, nvec in cases:
        n = sim.n[0]
        op = sim.op(name)
        rowptr, col, val = op.arrays()
        rg, _ = op.maps()
        n_global = sim.num_global("bfield")
        cuts = slab_cuts(rg, n_global, n + 1, world)
        r0, r1 = cuts[rank], cuts[rank + 1]
        lrp, lcol, lval = local_block(rowptr, rg[col], val, r0, r1)
        bmap = mx.MxMap(ctx, n_global, rg[r0:r1])
        for layout in (0, 1):
            A = mx.MxCrsMatrix.from_csr(bmap, bmap, lrp, lcol, lval, layout=layout)
            st = A.stats()
            assert st["ghosts"] > 0 and st["ghost_rows"] > 0, (label, st)
            x = mx.MxMultiVector(bmap, nvec, op.is_complex)
            y = mx.MxMultiVector(bmap, nvec, op.is_complex)
            x.random(4242)
            xg = np.empty((len(rg), nvec), dtype=x.dtype)
            for j in range(nvec):
                xg[:, j] = mx.hash_uniform(4242, rg, j, 0)
                if op.is_complex:
                    xg[:, j] += 1j * mx.hash_uniform(4242, rg, j, 1)
            assert np.array_equal(x.to_host(), xg[r0:r1]), "random is not P-invariant"
            for rep in range(3):                      # repeated applies reuse the ghost buffers
                A.apply(x, y)
            yg = op.apply(xg)
            got = y.to_host()
            if op.is_complex:
                err = np.linalg.norm(got - yg[r0:r1]) / np.linalg.norm(yg[r0:r1])
                assert err < 1e-14, (label, layout, err)
            else:
                assert np.array_equal(got, yg[r0:r1]), (label, layout, "halo apply differs from the global apply")
            np.testing.assert_allclose(x.norm2(), np.linalg.norm(xg, axis=0), rtol=1e-13)
            np.testing.assert_allclose(y.dot(x), np.einsum("ij,ij->j", xg.conj(), yg), rtol=1e-11, atol=1e-8)
            G = y.MvTransMv(1.0, x)
            np.testing.assert_allclose(G, xg.conj().T @ yg, rtol=1e-10, atol=1e-7)
            del A
        if rank == 0:
            print("case %s ok on %d ranks" % (label, world), flush=True)
    dist.barrier()
    print("RANK %d OK" % rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

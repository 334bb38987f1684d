"""Times the device operator assembly of the 256^3 pillbox and checks the assembled curl-curl against
the oracle's host-generated operator: same CSR arrays, same y = A x bits. Usage: python scripts/asm_bench.py [N] [--no-oracle]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maxwell_b200 as mx  # noqa: E402
from maxwell_b200 import assembly as asm  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
    with_oracle = "--no-oracle" not in sys.argv
    ctx = mx.Context(0)
    api = asm.gpu_api()
    out = {"n": n}

    def timed(label, fn):
        ctx.sync()
        t = time.time()
        r = fn()
        ctx.sync()
        out[label] = round(time.time() - t, 4)
        return r

    # warm the kernels once on a small grid
    w = asm.gpu_sim(ctx, 16, origin=(-0.5,) * 3, size=(1.0,) * 3)
    w.set_pec_shape(api.intersection([api.cylinder(0.4, (0, 0, 1), (0, 0, 0)), api.slab(0.8, (0, 0, 1), (0, 0, 0))])).setup()
    for name in ("curlCurl", "vecLapl", "scaLapl"):
        w.op(name)
    del w

    shape = api.intersection([api.cylinder(0.4, (0, 0, 1), (0, 0, 0)), api.slab(0.8, (0, 0, 1), (0, 0, 0))])
    sim = asm.gpu_sim(ctx, n, origin=(-0.5,) * 3, size=(1.0,) * 3)
    timed("fractions_s", lambda: sim.set_pec_shape(shape))
    timed("maps_s", lambda: sim.setup())
    out["dofs"] = {f: sim.map_size(f)[0] for f in asm.FIELDS}
    ops = {}
    for name in ("curlCurl", "gradDiv", "vecLapl", "scaLapl", "divB", "gradPsi", "dmA"):
        ops[name] = timed("op_%s_s" % name, lambda name=name: sim.op(name))
        out["nnz_" + name] = ops[name].nnz
    out["assembly_total_s"] = round(sum(v for k, v in out.items() if k.endswith("_s")), 4)
    bmap = timed("make_map_s", lambda: asm.make_map(sim, "bfield"))
    A = timed("layout_from_device_s", lambda: asm.to_crs(ops["curlCurl"], bmap, bmap))
    out["layout"] = A.stats()
    x = mx.MxMultiVector(bmap, 1)
    x.random(12345)
    y = mx.MxMultiVector(bmap, 1)
    A.apply(x, y)
    yh = y.to_host().reshape(-1)
    out["y_norm"] = float(np.linalg.norm(yh))
    if with_oracle:
        from oracle import oracle as orc
        t = time.time()
        o = orc.pillbox(n)
        out["oracle_setup_s"] = round(time.time() - t, 2)
        t = time.time()
        ref = o.op("curlCurl")
        out["oracle_curlCurl_s"] = round(time.time() - t, 2)
        same = all(np.array_equal(a, b) for a, b in zip(ref.arrays(), ops["curlCurl"].arrays()))
        out["curlCurl_csr_bit_exact"] = bool(same)
        out["maps_equal"] = bool(np.array_equal(o.map("bfield"), sim.map("bfield")))
        out["fractions_equal"] = bool(all(np.array_equal(o.full_fracs(f), sim.fracs(f)) for f in asm.FIELDS))
        out["apply_bit_exact"] = bool(np.array_equal(ref.apply(x.to_host().reshape(-1)), yh))
        out["host_threads"] = orc.lib().mxo_num_threads()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

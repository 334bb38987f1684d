"""Times the device assembly of config C4 (phc-192: sapphire sphere in a periodic cell, Bloch phases, complex128):
dielectric fractions, maps, invEps, curlCurl = curlE invEps curlB, vecLapl. Usage: python scripts/asm_bench_phc.py [N]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import maxwell_b200 as mx  # noqa: E402
from maxwell_b200 import assembly as asm  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    ctx = mx.Context(0)
    api = asm.gpu_api()
    asm.example_sim(ctx, "phc", 12, phase_shifts=(0.3, 0.2, 0.1)).op("vecLapl")       # load the kernels
    out = {"n": n}

    def timed(label, fn):
        ctx.sync()
        t = time.time()
        r = fn()
        ctx.sync()
        out[label] = round(time.time() - t, 4)
        return r

    sim = asm.gpu_sim(ctx, n, origin=(-0.5,) * 3, size=(1.0,) * 3, phase_shifts=(2.0, 1.0, 0.5))
    timed("dielectric_fractions_s", lambda: sim.add_dielectric(api.sphere(0.37, (0, 0, 0)), asm.SAPPHIRE))
    timed("maps_s", lambda: sim.setup())
    out["dofs"] = {f: sim.map_size(f)[0] for f in asm.FIELDS}
    cc = None
    for name in ("invEps", "curlCurl", "vecLapl"):
        m = timed("op_%s_s" % name, lambda name=name: sim.op(name))
        out["nnz_" + name] = m.nnz
        if name == "curlCurl":
            cc = m
    del m
    bmap = asm.make_map(sim, "bfield")
    A = timed("layout_s", lambda: asm.to_crs(cc, bmap, bmap))
    out["layout"] = A.stats()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Dev check: register-resident pipeline vs the wavefront on Cornell (same per-path values, throughput).
usage: gpu_resident.py [spp] [lib variants: name=BLOCK,MINBLOCKS ...]"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

SPP = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def run():
    from tuturenderer_b200 import api
    sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene")
    ctx = api.Context(0)
    small = sc.with_size(96, 96)
    ctx.upload(small)
    ctx.pipeline("wavefront")
    a = ctx.render_path(16, seed=3)
    sa = ctx.stats()
    ctx.pipeline("resident")
    b = ctx.render_path(16, seed=3)
    sb = ctx.stats()
    d = np.abs(a - b)
    print("96x96@16: max abs diff", d.max(), "rel", (d / (1e-6 + np.abs(a))).max(), "rays", sa["extend_rays"], sb["extend_rays"],
          sa["shadow_rays"], sb["shadow_rays"], "nan", sa["nan_samples"], sb["nan_samples"], flush=True)
    ctx.upload(sc.with_size(1024, 1024))
    for name in ("wavefront", "resident"):
        ctx.pipeline(name)
        ctx.render_path(16, seed=1)
        best = 1e9
        for rep in range(2):
            ctx.render_path(SPP, seed=2 + rep)
            best = min(best, ctx.stats()["gpu_ms"])
        st = ctx.stats()
        print(f"{name}: {1024 * 1024 * SPP / best * 1e-3:8.1f} Mpaths/s ({best:.1f} ms) trips {st['iterations']}", flush=True)


if os.environ.get("TUTU_CHILD"):
    run()
else:
    import subprocess
    subprocess.run([sys.executable, __file__, str(SPP)], env={**os.environ, "TUTU_CHILD": "1"})
    for v in sys.argv[2:]:
        name, cfg = v.split("=")
        lib = ROOT / "tuturenderer_b200" / f"libtutu_b200_{name}.so"
        if lib.exists():
            print("==", name, cfg, flush=True)
            subprocess.run([sys.executable, __file__, str(SPP)], env={**os.environ, "TUTU_CHILD": "1", "TUTU_LIB": str(lib)})

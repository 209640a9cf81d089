"""Dev sweep: wavefront lanes x paths in flight on Cornell 1024^2 @ 256 spp (device ms from TutuRenderStats)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api

sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(1024, 1024)
ctx = api.Context(0)
ctx.upload(sc)
SPP = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for lanes in (1, 2, 3):
    for pif in (8 << 20, 16 << 20):
        ctx.configure(pif, False, lanes)
        ctx.render_path(32, seed=1)
        best = 1e9
        for rep in range(2):
            ctx.render_path(SPP, seed=2 + rep)
            best = min(best, ctx.stats()["gpu_ms"])
        print(f"lanes {lanes} paths_in_flight {pif >> 20} Mi: {1024 * 1024 * SPP / best * 1e-3:8.1f} Mpaths/s ({best:.1f} ms)", flush=True)

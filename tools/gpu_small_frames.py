"""Dev: wavefront vs register-resident pipeline on small frames (device ms of one render)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api
sc0 = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene")
ctx = api.Context(0)
for (w, spp) in ((64, 16), (128, 16), (256, 16), (256, 64), (512, 16), (1024, 4)):
    ctx.upload(sc0.with_size(w, w))
    row = []
    for name in ("wavefront", "resident"):
        ctx.pipeline(name)
        ctx.render_path(spp, seed=1)
        best = 1e9
        for rep in range(3):
            ctx.render_path(spp, seed=2 + rep)
            best = min(best, ctx.stats()["gpu_ms"])
        row.append(best)
    print(f"{w}x{w} @ {spp} spp ({w * w * spp / 1e6:.2f} Mpaths): wavefront {row[0]:.3f} ms, resident {row[1]:.3f} ms", flush=True)

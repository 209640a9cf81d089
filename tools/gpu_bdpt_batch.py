"""Dev sweep: BDPT samples per batch on the Veach room 800x600 @ 32 spp."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api
sc = api.Scene.load(ROOT / "tests/golden/veach_80x60.tscene").with_size(800, 600)
ctx = api.Context(0)
ctx.upload(sc)
for cap in (1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    ctx.configure(cap, False, 0)
    ctx.render_bdpt(4, seed=1)
    ctx.render_bdpt(32, seed=2)
    st = ctx.stats()
    print(f"batch {cap >> 10} Ki samples: {800 * 600 * 32 / st['gpu_ms'] * 1e-3:.2f} Msamples/s ({st['gpu_ms']:.1f} ms, {st['kernel_launches']} launches)", flush=True)

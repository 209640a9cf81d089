"""Dev sweep on the configs[3] stand-in (general traversal + every shade path): lanes x capacity."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api

sc = api.Scene.load(ROOT / "tests/golden/glass_c4.tscene").with_size(1024, 1024)
ctx = api.Context(0)
ctx.upload(sc)
for SPP in (128,):
  for lanes in (1, 2):
    for pif in (16 << 20, 32 << 20):
        for prof in (False, True):
            ctx.configure(pif, prof, lanes)
            ctx.render_path(16, seed=1)
            ctx.render_path(SPP, seed=2)
            st = ctx.stats()
            print(f"spp {SPP} iters {st['iterations']} lanes {lanes} pif {pif >> 20} Mi prof {prof}: {1024 * 1024 * SPP / st['gpu_ms'] * 1e-3:8.1f} Mpaths/s ({st['gpu_ms']:.1f} ms) "
                  f"extend {st['extend_ms']:.0f} shade {st['shade_ms']:.0f} shadow {st['shadow_ms']:.0f} other {st['other_ms']:.0f}", flush=True)

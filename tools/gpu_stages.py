"""Dev: per-stage device ms of the wavefront on Cornell 1024^2 (CUDA events between the kernels)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api

sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(1024, 1024)
ctx = api.Context(0)
ctx.upload(sc)
SPP = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for lanes in (1, 2):
    for prof in (True, False):
        ctx.configure(0, prof, lanes)
        ctx.render_path(32, seed=1)
        ctx.render_path(SPP, seed=2)
        st = ctx.stats()
        print(f"lanes {lanes} profile {prof}: {1024 * 1024 * SPP / st['gpu_ms'] * 1e-3:8.1f} Mpaths/s  gpu_ms {st['gpu_ms']:.1f} extend {st['extend_ms']:.1f} "
              f"shade {st['shade_ms']:.1f} shadow {st['shadow_ms']:.1f} other {st['other_ms']:.1f} iters {st['iterations']} launches {st['kernel_launches']}", flush=True)

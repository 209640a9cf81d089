"""Development aid: one Cornell 1024x1024 render with per-stage timing; knobs come from the
environment (TUTU_LIB, TUTU_REFILL_MIN, TUTU_PIF, ...).  Prints one line."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = api.Context(0)
sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(1024, 1024)
ctx.upload(sc)
pif = int(os.environ.get("TUTU_PIF", "0"))
lanes = int(os.environ.get("TUTU_LANES", "0"))
ctx.configure(pif, False, lanes)
ctx.render_path(8, seed=1)
img = ctx.render_path(spp, seed=2)
st0 = ctx.stats()
ctx.configure(pif, True, lanes)
img = ctx.render_path(spp, seed=2)
st = ctx.stats()
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("TUTU_"))
print(f"[{tag}] {st0['paths'] / st0['gpu_ms'] * 1e-3:7.1f} Mpaths/s (unprofiled) | profiled {st['paths'] / st['gpu_ms'] * 1e-3:7.1f}: "
      f"extend {st['extend_ms']:.1f} shade {st['shade_ms']:.1f} shadow {st['shadow_ms']:.1f} other {st['other_ms']:.1f} ms; "
      f"iters {st['iterations']} mean {img.mean():.5f}")

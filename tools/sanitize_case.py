"""Smallest cases that touch every kernel family once (for one compute-sanitizer --tool memcheck pass)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api
G = ROOT / "tests" / "golden"
ctx = api.Context(0)
# small-scene kernels (Cornell)
sc = api.Scene.load(G / "cornell_256.tscene").with_size(24, 24)
ctx.upload(sc)
rays = np.fromfile(G / "cornell_rays.f32", np.float32).reshape(-1, 8)[:3000]
ctx.trace_closest(rays); ctx.trace_any(rays)
ctx.render_path(2, seed=1); ctx.render_bdpt(2, seed=1)
# general kernels: SAH tree + reference tree (irregular rays), binning threshold crossed, all materials / textures
hf = api.Scene.load(G / "hf24.tscene")
ctx.upload(hf)
r = api.synth_rays(0, 70000, seed=5)
r[::97, 4:7] = (0, -1, 0)
for mode in (0, 1, 3, 4, 6):  # production, literal, caller order, tree walk, compressed wide tree (experiment builds add 2, 10-17)
    ctx.set_traversal_mode(mode)
    ctx.trace_closest(r); ctx.trace_any(r)
ctx.set_traversal_mode(0)
mixed = api.Scene.load(G / "mixed.tscene").with_size(24, 24)
ctx.upload(mixed)
ctx.render_path(2, seed=2); ctx.render_bdpt(2, seed=2)
img = ctx.render_path(1, seed=3)
ctx.quantize(img)
print("sanitize case done")

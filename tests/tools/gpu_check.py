"""Quick on-GPU sanity run (development aid): parity of the CUDA path against the oracle on the
Cornell fixture plus a first throughput reading.  Not part of the test suite."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api
from oracle import oracle_py as O

s = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(128, 128)
osc = O.OracleScene(s)
ctx = api.Context(0)
ctx.upload(s)
i = ctx.info()
print("scene:", i.n_prims, i.n_nodes, i.depth, i.n_lights, i.device_bytes)
rays = osc.primary_rays()
rng = np.random.default_rng(0)
extra = np.zeros((20000, 8), np.float32)
extra[:, 0:3] = rng.uniform([0, 0, 0], [556, 548, 559], (20000, 3))
d = rng.normal(size=(20000, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
extra[:, 4:7] = d; extra[:, 7] = rng.uniform(10, 600, 20000)
rays = np.concatenate([rays, extra]).astype(np.float32)
for mode in (0, 1):
    ctx.set_traversal_mode(mode)
    hg = ctx.trace_closest(rays); ho = osc.trace_closest(rays)
    print(f"mode {mode} closest prim equal:", np.array_equal(hg["prim"], ho["prim"]), "bits equal:", np.array_equal(hg.view(np.uint8), ho.view(np.uint8)),
          "mismatch:", int((hg["prim"] != ho["prim"]).sum()))
    ag = ctx.trace_any(rays); ao = osc.trace_any(rays)
    print(f"mode {mode} any equal:", np.array_equal(ag, ao), int((ag != ao).sum()), ao.mean())
ctx.set_traversal_mode(0)
for spp in (1, 16):
    t = time.time(); g = ctx.render_path(spp, seed=7); tg = time.time() - t
    o = osc.render_path(spp, seed=7)
    diff = np.abs(g - o)
    print(f"spp {spp}: gpu mean {g.mean():.6f} oracle mean {o.mean():.6f} max|d| {diff.max():.4g} frac>1e-3 {(diff > 1e-3).mean():.5f} rmse {np.sqrt((diff**2).mean()):.5g}  host-s {tg:.3f}", ctx.stats())
s2 = s.with_size(1024, 1024)
ctx.upload(s2)
for spp in (4, 64):
    g = ctx.render_path(spp, seed=3)
    st = ctx.stats()
    print(f"1024^2 spp {spp}: {st['paths'] / st['gpu_ms'] * 1e-3:.1f} Mpaths/s, mean {g.mean():.5f}", st)
ctx.configure(0, True)
g = ctx.render_path(64, seed=3); st = ctx.stats()
print("profiled:", st)

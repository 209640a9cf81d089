"""Development aid (needs an experiment build: build.build_variant("exp", ["TUTU_EXPERIMENTS"]) and TUTU_LIB=<that .so>;
the shipped library rejects these modes): times the traversal flavours (tutu_set_traversal_mode 0 = persistent, 10/11/12 =
packet LOOP exact / LOOP FMNMX / rounds) on (a) Cornell bounce rays and (b) the height-field batches."""
import sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api
from oracle import oracle_py as O

stream = api.stream_handle(torch.cuda.current_stream().cuda_stream)


def timed(fn, d_rays, n, out, reps=5):
    fn(d_rays.data_ptr(), n, out.data_ptr(), stream); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(d_rays.data_ptr(), n, out.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(ctx, name, rays):
    n = len(rays)
    d_rays = torch.from_numpy(rays).cuda()
    d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    d_any = torch.empty(n, dtype=torch.uint8, device="cuda")
    ref = None
    for mode in (10, 11, 12, 0):
        ctx.set_traversal_mode(mode)
        mc = timed(ctx.trace_closest_device, d_rays, n, d_hits)
        ma = timed(ctx.trace_any_device, d_rays, n, d_any)
        cur = (d_hits.clone().view(torch.int32), d_any.clone())
        same = "" if ref is None else f" same-as-10: {bool((cur[0] == ref[0]).all())} {bool((cur[1] == ref[1]).all())}"
        ref = ref or cur
        print(f"{name:22s} mode {mode:2d}: closest {n / mc * 1e-3:8.1f} Mrays/s  any {n / ma * 1e-3:8.1f} Mrays/s{same}")
    ctx.set_traversal_mode(0)


ctx = api.Context(0)
sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(1024, 1024)
ctx.upload(sc)
osc = O.OracleScene(sc)
prim = osc.primary_rays()
hp = ctx.trace_closest(prim)
ok = hp["prim"] >= 0
pos = prim[ok, 0:3] + hp["t"][ok, None] * prim[ok, 4:7]
nrm = sc.prims["n"][hp["prim"][ok], 0:3]
rng = np.random.default_rng(1)
reps = 4
pos = np.repeat(pos, reps, 0); nrm = np.repeat(nrm, reps, 0)
d = rng.normal(size=pos.shape); d /= np.linalg.norm(d, axis=1, keepdims=True)
flip = (d * nrm).sum(1) < 0
d[flip] *= -1
bounce = np.zeros((len(pos), 8), np.float32)
bounce[:, 0:3] = pos + 5e-4 * nrm
bounce[:, 4:7] = d
bounce[:, 7] = rng.uniform(50, 600, len(pos))
prim[:, 7] = 900
run(ctx, "cornell primary (1M)", prim)
run(ctx, f"cornell bounce ({len(bounce) >> 20}M)", bounce)
prims = api.synth_heightfield(707)
ctx.upload(api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims)))
for kind in (0, 1):
    run(ctx, f"heightfield kind {kind} (8M)", api.synth_rays(kind, 1 << 23))

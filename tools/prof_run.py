"""Short, fixed workload for ncu (never a bench number): `render` = Cornell 1024x1024 @ 24 spp,
`rays` = 4 Mi rays of each synthetic kind against the 999 698-triangle height-field, `bdpt` = the
Veach room 800x600 @ 4 spp through tutu_render_bdpt, `build` = one upload of that height-field per device tree builder."""
import os
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api

what = sys.argv[1] if len(sys.argv) > 1 else "render"
ctx = api.Context(0)
if what == "render":
    sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(1024, 1024)
    ctx.upload(sc)
    import os
    ctx.configure(0, False, int(os.environ.get("TUTU_LANES", "0")))
    img = ctx.render_path(int(os.environ.get("TUTU_PROF_SPP", "24")), seed=5)
    print("render mean", float(img.mean()), ctx.stats())
elif what == "build":  # the traversal-tree builders on the 999 698-triangle scene (launch list only)
    prims = api.synth_heightfield(707)
    sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
    for builder in ("device_sah", "device_ploc", "device_lbvh"):
        ctx.builder(builder)
        ctx.upload(sc)
        print(builder, ctx.upload_stats())
elif what == "bdpt":
    sc = api.Scene.load(ROOT / "tests/golden/veach_80x60.tscene").with_size(800, 600)
    ctx.upload(sc)
    ctx.bdpt_queue_tracer(os.environ.get("TUTU_PROF_TRACER", "lanes"))  # what the library measures to be faster on this scene
    img = ctx.render_bdpt(4, seed=5)
    print("bdpt mean", float(img.mean()), ctx.stats())
else:
    G, N = 707, 1 << 22
    prims = api.synth_heightfield(G)
    sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
    ctx.upload(sc)
    for kind in (0, 1):
        rays = torch.from_numpy(api.synth_rays(kind, N)).cuda()
        hits = torch.empty((N, 4), dtype=torch.float32, device="cuda")
        blocked = torch.empty(N, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.trace_closest_device(rays.data_ptr(), N, hits.data_ptr())
        ctx.trace_any_device(rays.data_ptr(), N, blocked.data_ptr())
        torch.cuda.synchronize()
        print("kind", kind, "hit frac", float((hits.view(torch.int32)[:, 0] >= 0).float().mean()), "blocked", float(blocked.float().mean()))

"""Workload for the guard-band build (libtutu_b200_guard.so, -DTUTU_GUARDS; picked up through TUTU_LIB): every kernel
family and host loop of the library at sizes that fill and refill the queues, then the guard bands of every live
device allocation are checked (tutu_debug_guard_check).  compute-sanitizer is closed on the GPU pool; this is the
library's own detector for writes outside an allocation.  Prints one JSON line; tests/test_gpu_guards.py asserts on it."""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api
G = ROOT / "tests" / "golden"
checks = []
def check(tag):
    n, bad = api.guard_check()
    checks.append({"after": tag, "buffers": n, "bad_bytes": bad})

ctx = api.Context(0)
# small-scene kernels (Cornell): ray batches, the resident kernel, the wavefront with chunked queue reserves and refills
sc = api.Scene.load(G / "cornell_256.tscene")
ctx.upload(sc.with_size(24, 24))
rays = np.fromfile(G / "cornell_rays.f32", np.float32).reshape(-1, 8)
ctx.trace_closest(rays[:3001]); ctx.trace_any(rays[:3001]); ctx.trace_closest(rays[:1]); ctx.trace_any(rays[:33])
ctx.render_path(2, seed=1); ctx.render_bdpt(2, seed=1)
check("cornell 24x24")
ctx.upload(sc.with_size(333, 251))          # odd frame, the queue holds every path
for pipeline in ("wavefront", "resident", "auto"):
    ctx.pipeline(pipeline)
    ctx.render_path(7, seed=2)
ctx.pipeline("wavefront")
ctx.configure(1 << 16, False, 1)              # tiny queue: many refills, reserves at the very end of the arrays
ctx.render_path(9, seed=3)
ctx.configure(1 << 16, False, 2)
ctx.render_path(5, seed=4)
ctx.configure(0, False, 0)
ctx.pipeline("auto")
ctx.render_bdpt(3, seed=5)
check("cornell 333x251, small queues")
# tree kernels: every traversal mode and builder on the height-field (binning threshold crossed, irregular rays)
hf = api.Scene.load(G / "hf24.tscene")
r = api.synth_rays(0, 70001, seed=5)
r[::97, 4:7] = (0, -1, 0)
for builder in ("auto", "host_sah", "device_lbvh", "device_ploc", "device_sah"):
    try:
        ctx.builder(builder)
    except Exception:
        continue
    ctx.upload(hf)
    for mode in (0, 1, 3, 4, 6):
        ctx.set_traversal_mode(mode)
        ctx.trace_closest(r); ctx.trace_any(r)
    ctx.set_traversal_mode(0)
    for stack in ("shared", "local", "auto"):
        ctx.traversal_stack(stack)
        ctx.trace_closest(r); ctx.trace_any(r)
ctx.builder("auto")
check("height-field, all builders and modes")
# a bigger tree through the device builder, device-resident batch
import torch
prims = api.synth_heightfield(160)
big = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
ctx.upload(big)
n = (1 << 20) + 77
d_r = torch.from_numpy(api.synth_rays(1, n)).cuda()
d_h = torch.empty((n, 4), dtype=torch.float32, device="cuda"); d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ctx.trace_closest_device(d_r.data_ptr(), n, d_h.data_ptr()); ctx.trace_any_device(d_r.data_ptr(), n, d_a.data_ptr())
ctx.count_visits(d_r.data_ptr(), n, False)
h_host = ctx.trace_closest(d_r.cpu().numpy())  # host-pipelined chunks
check("51200-triangle tree, 2^20+77 rays")
# mixed materials / textures / spheres: class sort, general wavefront, BDPT; glass scene; Veach room
for name, size in (("mixed", (97, 61)), ("glass_c4", (160, 120)), ("veach_80x60", (80, 60))):
    s = api.Scene.load(G / f"{name}.tscene").with_size(*size)
    ctx.upload(s)
    ctx.render_path(5, seed=6)
    ctx.configure(1 << 15, False, 2)
    ctx.render_path(3, seed=7)
    ctx.configure(0, False, 0)
    ctx.render_bdpt(3, seed=8)
    for stack in ("shared", "local", "auto"):   # both stack flavours of the tree kernels (tutu_traversal_stack)
        ctx.traversal_stack(stack)
        ctx.render_path(2, seed=10)
        for tracer in ("packets", "lanes", "auto"):   # both BDPT queue tracers (tutu_bdpt_queue_tracer)
            ctx.bdpt_queue_tracer(tracer)
            ctx.render_bdpt(1, seed=11)
    for mode in (6, 0):
        ctx.set_traversal_mode(mode)
        ctx.render_path(2, seed=9)
    img = ctx.render_path(1, seed=3)
    ctx.quantize(img)
    ctx.postprocess(img, "hdr_bloom")
    check(name)
n0, bad0 = api.guard_check()
api._check(api.lib().tutu_debug_guard_poke(3))   # the detector must see a deliberate overrun
n1, bad1 = api.guard_check()
ctx.close()
print(json.dumps({"checks": checks, "buffers": n0, "bad_bytes": bad0, "bad_bytes_after_poke": bad1}))

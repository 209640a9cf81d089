"""Dev check: ray binning on/off on BASELINE.json configs[1] (parity bit for bit + Mrays/s)."""
import sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api

G = int(sys.argv[1]) if len(sys.argv) > 1 else 707
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
prims = api.synth_heightfield(G)
sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
ctx = api.Context(0)
ctx.upload(sc)
stream = api.stream_handle(torch.cuda.current_stream().cuda_stream)


def timed(fn, d_rays, n, out, reps=5):
    for _ in range(2):
        fn(d_rays.data_ptr(), n, out.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(d_rays.data_ptr(), n, out.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for kind in (0, 1):
    rays = api.synth_rays(kind, N)
    d_rays = torch.from_numpy(rays).cuda()
    res = {}
    MODES = [(3, "caller order"), (0, "binned")]  # an experiment build (TUTU_LIB, -DTUTU_EXPERIMENTS) also knows (16, "refill+bin"), (17, "leafbatch+bin")
    for mode, label in MODES:
        ctx.set_traversal_mode(mode)
        d_hits = torch.empty((N, 4), dtype=torch.float32, device="cuda")
        d_any = torch.empty(N, dtype=torch.uint8, device="cuda")
        ms_c = timed(ctx.trace_closest_device, d_rays, N, d_hits)
        ms_a = timed(ctx.trace_any_device, d_rays, N, d_any)
        res[mode] = (d_hits.view(torch.int32).clone(), d_any.clone())
        print(f"kind {kind} {label:15s}: closest {N / ms_c * 1e-3:8.1f} Mrays/s ({ms_c:6.2f} ms)   any {N / ms_a * 1e-3:8.1f} Mrays/s ({ms_a:6.2f} ms)")
    for mode, label in MODES[1:]:
        print(f"   {label}: closest identical {bool((res[3][0] == res[mode][0]).all())}  any identical {bool((res[3][1] == res[mode][1]).all())}")

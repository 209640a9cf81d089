"""configs[1] any-hit batches under the experiment build's traversal modes (TUTU_LIB = a -DTUTU_EXPERIMENTS library):
0 = production (per-lane structured walk), 15 = shared-memory stack, 17 = warp walk with batched leaf tests."""
import sys, json
sys.path.insert(0, '/root/repo')
import torch
from tuturenderer_b200 import api
prims = api.synth_heightfield(707)
sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
ctx = api.Context(0); ctx.upload(sc)
N = 1 << 24
stream = api.stream_handle(torch.cuda.current_stream().cuda_stream)
for kind in (0, 1):
    rays = torch.from_numpy(api.synth_rays(kind, N)).cuda()
    ref = None
    for mode in (0, 15, 17, 0):
        ctx.set_traversal_mode(mode)
        a = torch.empty(N, dtype=torch.uint8, device='cuda')
        for _ in range(2):
            ctx.trace_any_device(rays.data_ptr(), N, a.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ctx.trace_any_device(rays.data_ptr(), N, a.data_ptr(), stream)
        e1.record(); torch.cuda.synchronize()
        if ref is None:
            ref = a
        print(json.dumps({'kind': kind, 'mode': mode, 'any_mrays_s': round(N * 5 / e0.elapsed_time(e1) * 1e-3, 1), 'same': bool((a == ref).all())}), flush=True)
    del rays

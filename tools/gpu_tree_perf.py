"""Throughput of the tree-walking workloads for the library picked by TUTU_LIB (experiment builds): configs[1] ray
batches (closest / any, both ray kinds), the glass / texture scene, the Veach room BDPT.  One line per workload."""
import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from tuturenderer_b200 import api
G = '/root/repo/tests/golden/'
tag = os.environ.get('TUTU_LIB', 'default').split('libtutu_b200')[-1]
out = {'lib': tag}
prims = api.synth_heightfield(707)
sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
ctx = api.Context(0)
ctx.upload(sc)
N = 1 << 24
stream = api.stream_handle(torch.cuda.current_stream().cuda_stream)
for kind in (0, 1):
    rays = torch.from_numpy(api.synth_rays(kind, N)).cuda()
    h = torch.empty((N, 4), dtype=torch.float32, device='cuda')
    a = torch.empty(N, dtype=torch.uint8, device='cuda')
    for name, fn, dst in (('closest', ctx.trace_closest_device, h), ('any', ctx.trace_any_device, a)):
        for _ in range(2):
            fn(rays.data_ptr(), N, dst.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn(rays.data_ptr(), N, dst.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        out[f'k{kind}_{name}'] = round(N * 5 / e0.elapsed_time(e1) * 1e-3, 1)
    out[f'k{kind}_sum'] = int(h.view(torch.int32)[:, 0].long().sum().item()) ^ int(a.long().sum().item())
    del rays, h, a
ctx.close()
sc = api.Scene.load(G + 'glass_c4.tscene').with_size(1024, 1024)
ctx = api.Context(0); ctx.upload(sc)
for k in range(2):
    ctx.render_path(32, seed=k)
ctx.render_path(128, seed=9)
out['glass'] = round(1024 * 1024 * 128 / ctx.stats()['gpu_ms'] * 1e-3, 1)
ctx.close()
sc = api.Scene.load(G + 'veach_80x60.tscene').with_size(800, 600)
ctx = api.Context(0); ctx.upload(sc)
for k in range(2):
    ctx.render_bdpt(16, seed=k)
ctx.render_bdpt(128, seed=9)
out['bdpt'] = round(800 * 600 * 128 / ctx.stats()['gpu_ms'] * 1e-3, 1)
ctx.close()
print(json.dumps(out), flush=True)

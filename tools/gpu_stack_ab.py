"""Traversal stack in shared memory vs in local memory (tutu_traversal_stack) as a
function of the tree size: any-hit batches (both ray kinds) over height-fields of G x G quads, the glass scene, the Veach room."""
import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from tuturenderer_b200 import api
G_ = '/root/repo/tests/golden/'
stream = api.stream_handle(torch.cuda.current_stream().cuda_stream)
N = 1 << 23
for G in (24, 64, 128, 200, 280, 400, 707):
    prims = api.synth_heightfield(G)
    sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
    rec = {'G': G, 'tris': len(prims), 'tree_MB': round(len(prims) * (2 * 64 + 48) / 2**20, 2)}
    for shared in (0, 1):
        ctx = api.Context(0); ctx.traversal_stack('shared' if shared else 'local'); ctx.upload(sc)
        for kind in (0, 1):
            rays = torch.from_numpy(api.synth_rays(kind, N)).cuda()
            a = torch.empty(N, dtype=torch.uint8, device='cuda')
            for _ in range(2):
                ctx.trace_any_device(rays.data_ptr(), N, a.data_ptr(), stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                ctx.trace_any_device(rays.data_ptr(), N, a.data_ptr(), stream)
            e1.record(); torch.cuda.synchronize()
            rec[f'any_k{kind}_shared{shared}'] = round(N * 4 / e0.elapsed_time(e1) * 1e-3)
            del rays
        ctx.close()
    print(json.dumps(rec), flush=True)
for shared in (0, 1):
    rec = {'shared': shared}
    sc = api.Scene.load(G_ + 'glass_c4.tscene').with_size(1024, 1024)
    ctx = api.Context(0); ctx.traversal_stack('shared' if shared else 'local'); ctx.upload(sc)
    for k in range(2):
        ctx.render_path(32, seed=k)
    img = ctx.render_path(128, seed=9)
    rec['glass'] = round(1024 * 1024 * 128 / ctx.stats()['gpu_ms'] * 1e-3, 1); rec['glass_mean'] = round(float(img.mean()), 6)
    ctx.close()
    sc = api.Scene.load(G_ + 'veach_80x60.tscene').with_size(800, 600)
    ctx = api.Context(0); ctx.traversal_stack('shared' if shared else 'local'); ctx.bdpt_queue_tracer('packets'); ctx.upload(sc)
    for k in range(2):
        ctx.render_bdpt(16, seed=k)
    im = ctx.render_bdpt(128, seed=9)
    rec['bdpt'] = round(800 * 600 * 128 / ctx.stats()['gpu_ms'] * 1e-3, 1); rec['bdpt_mean'] = round(float(im.mean()), 6)
    ctx.close()
    print(json.dumps(rec), flush=True)

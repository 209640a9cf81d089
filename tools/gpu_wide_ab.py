"""A/B of the compressed 8-wide tree (traversal mode 6) against the binary SAH tree (mode 0) on a B200:
configs[1] ray batches (both ray kinds, closest + any hit, visits per ray), the glass / texture scene and the
Veach room BDPT.  Results of both modes are compared bit for bit."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from tuturenderer_b200 import api

G = '/root/repo/tests/golden/'
out = {}
prims = api.synth_heightfield(707)
sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
ctx = api.Context(0)
ctx.upload(sc)
N = 1 << 24
stream = api.stream_handle(torch.cuda.current_stream().cuda_stream)
for kind in (0, 1):
    rays = torch.from_numpy(api.synth_rays(kind, N)).cuda()
    hits = {m: torch.empty((N, 4), dtype=torch.float32, device='cuda') for m in (6, 0)}
    anys = {m: torch.empty(N, dtype=torch.uint8, device='cuda') for m in (6, 0)}
    for mode in (6, 0):
        ctx.set_traversal_mode(mode)
        info = ctx.info()
        rec = {'width': info.trav_width, 'nodes': info.trav_nodes, 'depth': info.trav_depth}
        for name, fn, dst in (('closest', ctx.trace_closest_device, hits[mode]), ('any', ctx.trace_any_device, anys[mode])):
            for _ in range(2):
                fn(rays.data_ptr(), N, dst.data_ptr(), stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn(rays.data_ptr(), N, dst.data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            rec[name + '_mrays_s'] = N * 5 / e0.elapsed_time(e1) * 1e-3
        nc, pc = ctx.count_visits(rays.data_ptr(), N, False)
        na, pa = ctx.count_visits(rays.data_ptr(), N, True)
        rec.update(nodes_per_ray=nc / N, prims_per_ray=pc / N, any_nodes_per_ray=na / N, any_prims_per_ray=pa / N)
        out[f'rays_kind{kind}_mode{mode}'] = rec
        print(f'rays kind {kind} mode {mode}:', json.dumps(rec), flush=True)
    same_h = bool((hits[6].view(torch.int32) == hits[0].view(torch.int32)).all())
    same_a = bool((anys[6] == anys[0]).all())
    print(f'rays kind {kind}: wide == binary: closest {same_h}, any {same_a}', flush=True)
    out[f'rays_kind{kind}_identical'] = [same_h, same_a]
    del rays, hits, anys
ctx.close()

# glass / texture scene (configs[3] stand-in), 1024^2
sc = api.Scene.load(G + 'glass_c4.tscene').with_size(1024, 1024)
ctx = api.Context(0)
ctx.upload(sc)
img = {}
for mode in (6, 0):
    ctx.set_traversal_mode(mode)
    for k in range(2):
        ctx.render_path(32, seed=k)
    img[mode] = ctx.render_path(128, seed=9)
    st = ctx.stats()
    out[f'glass_mode{mode}'] = {'mpaths_s': 1024 * 1024 * 128 / st['gpu_ms'] * 1e-3, 'gpu_ms': st['gpu_ms']}
    print(f'glass mode {mode}:', out[f'glass_mode{mode}'], flush=True)
    ctx.configure(0, True, 1)
    ctx.render_path(64, seed=3)
    st = ctx.stats()
    print('   stages (1 lane, 64 spp): extend %.1f shade %.1f shadow %.1f other %.1f ms' % (st['extend_ms'], st['shade_ms'], st['shadow_ms'], st['other_ms']), flush=True)
    ctx.configure(0, False, 0)
d = np.abs(img[6] - img[0])
print('glass: max abs diff between modes', float(d.max()), 'pixels differing > 1e-3 rel', float((d > 1e-3 * (1 + np.abs(img[0]))).any(-1).mean()))
ctx.close()

# Veach room BDPT 800x600
sc = api.Scene.load(G + 'veach_80x60.tscene').with_size(800, 600)
ctx = api.Context(0)
ctx.upload(sc)
for mode in (6, 0):
    ctx.set_traversal_mode(mode)
    for k in range(2):
        ctx.render_bdpt(16, seed=k)
    im = ctx.render_bdpt(128, seed=9)
    st = ctx.stats()
    out[f'bdpt_mode{mode}'] = {'msamples_s': 800 * 600 * 128 / st['gpu_ms'] * 1e-3, 'gpu_ms': st['gpu_ms'], 'mean': float(im.mean())}
    print(f'bdpt mode {mode}:', out[f'bdpt_mode{mode}'], flush=True)
ctx.close()
json.dump(out, open('/root/repo/gpurun_out/r02b_wide_ab.json', 'w'), indent=1)

"""Traversal tree built on the GPU (linear BVH) against the host's binned-SAH tree: upload time breakdown,
configs[1] ray throughput and visits per ray, glass scene and Veach BDPT throughput."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from tuturenderer_b200 import api

G = '/root/repo/tests/golden/'
prims = api.synth_heightfield(707)
t = time.perf_counter()
nodes = api.bvh_build(prims)
print('reference-topology midpoint build (host): %.1f ms' % ((time.perf_counter() - t) * 1e3))
sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=nodes)
ctx = api.Context(0)
N = 1 << 24
stream = api.stream_handle(torch.cuda.current_stream().cuda_stream)
rays = {k: torch.from_numpy(api.synth_rays(k, N)).cuda() for k in (0, 1)}
hits = {}
for builder in ('host_sah', 'device_lbvh', 'device_ploc', 'device_sah'):
    ctx.builder(builder)
    ctx.upload(sc)
    for rep in range(2):
        t = time.perf_counter()
        ctx.upload(sc)
        wall = (time.perf_counter() - t) * 1e3
    print(builder, 'upload wall %.1f ms' % wall, json.dumps(ctx.upload_stats()), flush=True)
    for kind in (0, 1):
        h = torch.empty((N, 4), dtype=torch.float32, device='cuda')
        a = torch.empty(N, dtype=torch.uint8, device='cuda')
        rec = {}
        for name, fn, dst in (('closest', ctx.trace_closest_device, h), ('any', ctx.trace_any_device, a)):
            for _ in range(2):
                fn(rays[kind].data_ptr(), N, dst.data_ptr(), stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn(rays[kind].data_ptr(), N, dst.data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            rec[name] = round(N * 5 / e0.elapsed_time(e1) * 1e-3, 1)
        nc, pc = ctx.count_visits(rays[kind].data_ptr(), N, False)
        rec.update(nodes_per_ray=round(nc / N, 2), prims_per_ray=round(pc / N, 2))
        print(' ', builder, 'kind', kind, rec, flush=True)
        hits[(builder, kind)] = (h, a)
for kind in (0, 1):
    for dev in ('device_lbvh', 'device_ploc', 'device_sah'):
        same = bool((hits[('host_sah', kind)][0].view(torch.int32) == hits[(dev, kind)][0].view(torch.int32)).all()) and \
            bool((hits[('host_sah', kind)][1] == hits[(dev, kind)][1]).all())
        print('kind', kind, dev, '== host-built:', same)
ctx.close()
del rays, hits

for name, file, w, h, kind in (('glass', 'glass_c4.tscene', 1024, 1024, 'pt'), ('veach', 'veach_80x60.tscene', 800, 600, 'bdpt')):
    sc = api.Scene.load(G + file).with_size(w, h)
    ctx = api.Context(0)
    for builder in ('host_sah', 'device_lbvh', 'device_ploc', 'device_sah'):
        ctx.builder(builder)
        ctx.upload(sc)
        render = ctx.render_path if kind == 'pt' else ctx.render_bdpt
        for k in range(2):
            render(32 if kind == 'pt' else 16, seed=k)
        render(128, seed=9)
        st = ctx.stats()
        print(name, builder, json.dumps(ctx.upload_stats()), '%.1f M/s' % (w * h * 128 / st['gpu_ms'] * 1e-3), flush=True)
    ctx.close()

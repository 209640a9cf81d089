"""What would ordering the wavefront / BDPT queues buy on scenes that are a room around small detailed objects?
Bounce-like rays (origins = first hits of random rays from the eye, random directions into the room) on the glass /
texture scene and the Veach room, traced by the batch kernels in the caller's (random) order (mode 3) and in the
binned order (mode 0, time includes the counting sort).  Prints Mrays/s and visits per ray."""
import sys, json
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from tuturenderer_b200 import api
G = '/root/repo/tests/golden/'
N = 1 << 23
stream = api.stream_handle(torch.cuda.current_stream().cuda_stream)
def unit(v):
    return v / v.norm(dim=1, keepdim=True)
for name in ('glass_c4', 'veach_80x60'):
    sc = api.Scene.load(G + name + '.tscene')
    ctx = api.Context(0); ctx.upload(sc)
    g = torch.Generator(device='cuda'); g.manual_seed(7)
    eye = torch.tensor(sc.eye, dtype=torch.float32, device='cuda')
    view = torch.tensor(sc.viewdir, dtype=torch.float32, device='cuda')
    d0 = unit(unit(view[None, :]) + 0.45 * torch.randn((N, 3), device='cuda', generator=g))
    rays = torch.zeros((N, 8), dtype=torch.float32, device='cuda')
    rays[:, 0:3] = eye; rays[:, 4:7] = d0
    hits = torch.empty((N, 4), dtype=torch.float32, device='cuda')
    torch.cuda.synchronize()
    ctx.trace_closest_device(rays.data_ptr(), N, hits.data_ptr(), stream)
    torch.cuda.synchronize()
    ok = hits[:, 0].view(torch.int32) >= 0
    t = hits[:, 1:2]
    p = eye[None, :] + t * d0
    d1 = unit(torch.randn((N, 3), device='cuda', generator=g))
    d1 = torch.where(((d1 * d0).sum(1, keepdim=True) > 0), -d1, d1)      # back into the room
    sec = torch.zeros((N, 8), dtype=torch.float32, device='cuda')
    sec[:, 0:3] = p + 1e-3 * d1; sec[:, 4:7] = d1
    sec = sec[ok].contiguous()
    M = sec.shape[0]
    out = {'scene': name, 'rays': M}
    res = {}
    for mode, tag in ((3, 'caller_order'), (0, 'binned')):
        ctx.set_traversal_mode(mode)
        h = torch.empty((M, 4), dtype=torch.float32, device='cuda')
        for _ in range(2):
            ctx.trace_closest_device(sec.data_ptr(), M, h.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ctx.trace_closest_device(sec.data_ptr(), M, h.data_ptr(), stream)
        e1.record(); torch.cuda.synchronize()
        out[tag + '_mrays_s'] = round(M * 5 / e0.elapsed_time(e1) * 1e-3, 1)
        res[mode] = h
    nc, pc = ctx.count_visits(sec.data_ptr(), M, False)
    out['nodes_per_ray'] = round(nc / M, 2); out['prims_per_ray'] = round(pc / M, 2)
    out['identical'] = bool((res[0].view(torch.int32) == res[3].view(torch.int32)).all())
    print(json.dumps(out), flush=True)
    ctx.close()

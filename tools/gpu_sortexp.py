"""Experiment: how much does a FINE sort of the ray batch buy the pruned walk? (rays reordered with torch)"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api

G, N = 707, 1 << 24
prims = api.synth_heightfield(G)
sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
ctx = api.Context(0)
ctx.upload(sc)
ctx.set_traversal_mode(3)
stream = api.stream_handle(torch.cuda.current_stream().cuda_stream)


def timed(fn, d_rays, n, out, reps=3):
    fn(d_rays.data_ptr(), n, out.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(d_rays.data_ptr(), n, out.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def morton2(x, z, bits):
    k = torch.zeros_like(x)
    for b in range(bits):
        k |= ((x >> b) & 1) << (2 * b)
        k |= ((z >> b) & 1) << (2 * b + 1)
    return k


def morton3(x, y, z, bits):
    k = torch.zeros_like(x)
    for b in range(bits):
        k |= ((x >> b) & 1) << (3 * b)
        k |= ((y >> b) & 1) << (3 * b + 1)
        k |= ((z >> b) & 1) << (3 * b + 2)
    return k


for kind in (0, 1):
    rays = torch.from_numpy(api.synth_rays(kind, N)).cuda()
    o, d = rays[:, 0:3], rays[:, 4:7]
    lo = torch.tensor([0.0, -0.25, 0.0], device="cuda"); hi = torch.tensor([1.0, 0.3, 1.0], device="cuda")
    inv = 1.0 / d
    t0, t1 = (lo - o) * inv, (hi - o) * inv
    te = torch.minimum(t0, t1).max(dim=1).values.clamp(min=0)
    p = o + te[:, None] * d
    u = ((p - lo) / (hi - lo)).clamp(0, 0.999999)
    l1 = d.abs().sum(1)
    du, dv = d[:, 0] / l1, d[:, 2] / l1
    neg = d[:, 1] < 0
    du2 = torch.where(neg, (1 - dv.abs()) * torch.sign(du), du)
    dv2 = torch.where(neg, (1 - du.abs()) * torch.sign(dv), dv)
    du, dv = (du2 * 0.5 + 0.5).clamp(0, 0.999999), (dv2 * 0.5 + 0.5).clamp(0, 0.999999)
    variants = {}
    variants["unsorted"] = None
    for ob, db in ((5, 3), (7, 0), (7, 2), (7, 3), (9, 0), (9, 2), (10, 3), (6, 4)):
        q = (u * (1 << ob)).long()
        cell = morton3(q[:, 0], q[:, 1], q[:, 2], ob)
        dirbin = ((du * (1 << db)).long() << db) | (dv * (1 << db)).long() if db else torch.zeros_like(cell)
        variants[f"dir{db}-major/origin{ob}"] = (dirbin << (3 * ob)) | cell
        if db:
            variants[f"origin{ob}-major/dir{db}"] = (cell << (2 * db)) | dirbin
    for name, key in variants.items():
        r = rays if key is None else rays[torch.argsort(key)].contiguous()
        hits = torch.empty((N, 4), dtype=torch.float32, device="cuda")
        anyb = torch.empty(N, dtype=torch.uint8, device="cuda")
        ms_c = timed(ctx.trace_closest_device, r, N, hits)
        ms_a = timed(ctx.trace_any_device, r, N, anyb)
        print(f"kind {kind} {name:26s}: closest {N / ms_c * 1e-3:8.1f} Mrays/s ({ms_c:6.2f} ms)  any {N / ms_a * 1e-3:8.1f} Mrays/s ({ms_a:6.2f} ms)", flush=True)
        del r

"""Development aid: BASELINE.json configs[1] on the GPU — G x G height-field (2*G*G triangles),
N synthetic rays; pruned vs literal parity at full size, visit counts, Mrays/s, slack sweep."""
import os, sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api

G = int(sys.argv[1]) if len(sys.argv) > 1 else 707
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
SWEEP = "--sweep" in sys.argv
t = time.time(); prims = api.synth_heightfield(G); print("mesh", len(prims), f"{time.time()-t:.2f}s")
t = time.time(); nodes = api.bvh_build(prims); print("bvh", len(nodes), f"{time.time()-t:.2f}s")
sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=nodes)
ctx = api.Context(0)
t = time.time(); ctx.upload(sc); print("upload", f"{time.time()-t:.2f}s", "depth", ctx.info().depth, "bytes", ctx.info().device_bytes)
stream = api.stream_handle(torch.cuda.current_stream().cuda_stream)


def timed(fn, d_rays, n, out, reps=3):
    fn(d_rays.data_ptr(), n, out.data_ptr(), stream); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(d_rays.data_ptr(), n, out.data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for kind in (0, 1):
    t = time.time(); rays = api.synth_rays(kind, N); print(f"rays kind {kind}", f"{time.time()-t:.2f}s")
    d_rays = torch.from_numpy(rays).cuda()
    d_hits = torch.empty((N, 4), dtype=torch.float32, device="cuda")
    d_any = torch.empty(N, dtype=torch.uint8, device="cuda")
    # literal walk = the reference's recursion; full batch
    ctx.set_traversal_mode(1)
    ms = timed(ctx.trace_closest_device, d_rays, N, d_hits, reps=1); lit_c = d_hits.clone()
    print(f"kind {kind} literal closest: {N / ms * 1e-3:.1f} Mrays/s ({ms:.1f} ms)")
    ms = timed(ctx.trace_any_device, d_rays, N, d_any, reps=1); lit_a = d_any.clone()
    print(f"kind {kind} literal any:     {N / ms * 1e-3:.1f} Mrays/s ({ms:.1f} ms)")
    ctx.set_traversal_mode(0)
    settings = [None]
    if SWEEP:
        settings = [("0", "0"), ("0.0009765625", None), ("0.0078125", None), ("0.015625", None), ("0.03125", None)]
    for st in settings:
        if st is not None:
            os.environ["TUTU_PRUNE_REL"] = st[0]
            if st[1] is not None:
                os.environ["TUTU_PRUNE_ABS"] = st[1]
            else:
                os.environ.pop("TUTU_PRUNE_ABS", None)
            ctx.upload(sc)
        ms_c = timed(ctx.trace_closest_device, d_rays, N, d_hits)
        ms_a = timed(ctx.trace_any_device, d_rays, N, d_any)
        a, b = d_hits.view(torch.int32), lit_c.view(torch.int32)
        bad = (a != b).any(1)
        nn, pp = ctx.count_visits(d_rays.data_ptr(), N, False)
        na, pa = ctx.count_visits(d_rays.data_ptr(), N, True)
        print(f"kind {kind} slack {st}: closest {N / ms_c * 1e-3:.1f} Mrays/s ({ms_c:.2f} ms) any {N / ms_a * 1e-3:.1f} Mrays/s ({ms_a:.2f} ms) | "
              f"closest mismatches vs literal {int(bad.sum())}/{N}, any mismatches {int((d_any != lit_a).sum())} | "
              f"nodes/ray {nn / N:.2f} prims/ray {pp / N:.2f} bytes/ray {48 + 64 * nn / N + 48 * pp / N:.0f}; any nodes/ray {na / N:.2f} prims/ray {pa / N:.2f}")
        if int(bad.sum()):
            idx = torch.nonzero(bad)[:8, 0]
            for i in idx.tolist():
                print("   ray", i, rays[i].tolist(), "pruned", a[i, 0].item(), d_hits[i, 1:].tolist(), "literal", b[i, 0].item(), lit_c[i, 1:].tolist())
    print(f"  hit frac {float((lit_c.view(torch.int32)[:, 0] >= 0).float().mean()):.4f} blocked frac {float(lit_a.float().mean()):.4f}")

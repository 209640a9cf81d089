"""Development aid: BASELINE.json configs[1] on the GPU — G x G height-field (2*G*G triangles),
N synthetic rays; pruned vs unpruned parity at full size, visit counts, Mrays/s."""
import sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api

G = int(sys.argv[1]) if len(sys.argv) > 1 else 707
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
t = time.time(); prims = api.synth_heightfield(G); print("mesh", len(prims), f"{time.time()-t:.2f}s")
t = time.time(); nodes = api.bvh_build(prims); print("bvh", len(nodes), f"{time.time()-t:.2f}s")
sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=nodes)
ctx = api.Context(0)
t = time.time(); ctx.upload(sc); print("upload", f"{time.time()-t:.2f}s", "depth", ctx.info().depth, "bytes", ctx.info().device_bytes)
for kind in (0, 1):
    t = time.time(); rays = api.synth_rays(kind, N); print(f"rays kind {kind}", f"{time.time()-t:.2f}s")
    d_rays = torch.from_numpy(rays).cuda()
    d_hits = torch.empty((N, 4), dtype=torch.float32, device="cuda")
    d_any = torch.empty(N, dtype=torch.uint8, device="cuda")
    res = {}
    for mode in (0, 1):
        if mode == 1 and N > (1 << 22):
            n = 1 << 22  # the literal walk is ~10x slower; check a 4M prefix
        else:
            n = N
        ctx.set_traversal_mode(mode)
        for name, fn, out in (("closest", ctx.trace_closest_device, d_hits), ("any", ctx.trace_any_device, d_any)):
            fn(d_rays.data_ptr(), n, out.data_ptr()); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s = torch.cuda.current_stream()
            e0.record(); fn(d_rays.data_ptr(), n, out.data_ptr(), s.cuda_stream); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            res[(mode, name)] = out[:n].clone()
            print(f"kind {kind} mode {mode} {name}: {n / ms * 1e-3:.1f} Mrays/s ({ms:.2f} ms)")
    n = min(N, 1 << 22)
    a, b = res[(0, "closest")][:n].view(torch.int32), res[(1, "closest")][:n].view(torch.int32)
    print("  pruned == literal (closest bits):", bool((a == b).all()), "mismatch rays:", int((a != b).any(1).sum()),
          "hit frac", float((a[:, 0] >= 0).float().mean()))
    print("  pruned == literal (any):", bool((res[(0, 'any')][:n] == res[(1, 'any')][:n]).all()), float(res[(0, 'any')].float().mean()))
    ctx.set_traversal_mode(0)
    for any_hit in (False, True):
        nn, pp = ctx.count_visits(d_rays.data_ptr(), N, any_hit)
        print(f"  visits any={any_hit}: nodes/ray {nn / N:.2f} prims/ray {pp / N:.2f} -> bytes/ray {32 + (1 if any_hit else 16) + 64 * nn / N + 48 * pp / N:.0f}")

"""Dev: (a) upload cost per scene, (b) host-buffer ray batches (tutu_trace_closest) for the pipeline knobs."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tuturenderer_b200 import api

if len(sys.argv) > 1 and sys.argv[1] == "upload":
    ctx = api.Context(0)
    for name in ("cornell_256", "glass_c4", "veach_80x60"):
        sc = api.Scene.load(ROOT / f"tests/golden/{name}.tscene").with_size(1024, 1024)
        host = torch.empty(1024 * 1024 * 3, dtype=torch.float32, pin_memory=True)
        ctx.upload(sc)
        ctx.render_path_ptr(16, 1, host.data_ptr())
        for rep in range(2):
            t0 = time.perf_counter(); ctx.upload(sc); t1 = time.perf_counter()
            ctx.render_path_ptr(16, 2, host.data_ptr()); t2 = time.perf_counter()
            print(name, f"upload {1e3 * (t1 - t0):.1f} ms, render16 wall {1e3 * (t2 - t1):.1f} ms, gpu_ms {ctx.stats()['gpu_ms']:.1f}", flush=True)
else:
    G, N = 707, 1 << 24
    prims = api.synth_heightfield(G)
    sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
    ctx = api.Context(0)
    ctx.upload(sc)
    for kind in (0, 1):
        h_rays = torch.empty((N, 8), dtype=torch.float32, pin_memory=True)
        api.synth_rays(kind, N, out=h_rays.numpy())
        h_hits = torch.empty((N, 4), dtype=torch.float32, pin_memory=True)
        for _ in range(2):
            ctx.trace_closest_ptr(h_rays.data_ptr(), N, h_hits.data_ptr())
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.trace_closest_ptr(h_rays.data_ptr(), N, h_hits.data_ptr())
        dt = (time.perf_counter() - t0) / 3
        print(os.environ.get("TUTU_HOST_SLOTS"), os.environ.get("TUTU_HOST_CHUNK_LOG2"), "kind", kind, f"{N / dt * 1e-6:.0f} Mrays/s e2e ({dt * 1e3:.2f} ms)", flush=True)

"""Dev check on the GPU box: tutu_render_bdpt vs the CPU oracle on the same Philox stream, plus a
throughput figure.  Not a test, not a bench number."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

G = ROOT / "tests" / "golden"


def compare(name, sc, spp, seed=3):
    ctx = api.Context(0)
    ctx.upload(sc)
    g = ctx.render_bdpt(spp, seed=seed)
    st = ctx.stats()
    o, cnt = O.OracleScene(sc).render_bdpt(spp, seed=seed, counters=True)
    d = np.abs(g - o)
    bad = (d > 1e-3 * (1 + np.abs(o))).any(-1)
    print(f"{name}: {sc.width}x{sc.height}@{spp}  gpu mean {g.mean((0, 1))}  oracle mean {o.mean((0, 1))}")
    print(f"   pixels differing {bad.mean():.4%}  median |d| {np.median(d):.3g}  finite {np.isfinite(g).all()} / {np.isfinite(o).all()}")
    print(f"   gpu extend {st['extend_rays']} shadow {st['shadow_rays']} | oracle closest {cnt[0]} any {cnt[1]} connections {cnt[2]}"
          f" | {st['gpu_ms']:.2f} ms, {st['kernel_launches']} launches")
    if bad.any():
        ys, xs = np.nonzero(bad)
        for y, x in list(zip(ys, xs))[:5]:
            print("   ", y, x, g[y, x], o[y, x])
    ctx.close()


def main():
    cornell = api.Scene.load(G / "cornell_256.tscene")
    compare("cornell", cornell.with_size(48, 48), 8)
    compare("mixed", api.Scene.load(G / "mixed.tscene").with_size(48, 48), 8)
    if (G / "veach_80x60.tscene").exists():
        veach = api.Scene.load(G / "veach_80x60.tscene")
        compare("veach", veach.with_size(40, 30), 8)
        # throughput
        ctx = api.Context(0)
        for sc, spp, label in ((veach.with_size(800, 600), 16, "veach 800x600"), (cornell.with_size(1024, 1024), 8, "cornell 1024^2")):
            ctx.upload(sc)
            ctx.render_bdpt(2, seed=1)
            t = time.perf_counter()
            img = ctx.render_bdpt(spp, seed=2)
            dt = time.perf_counter() - t
            st = ctx.stats()
            print(f"{label} @ {spp} spp: {dt * 1e3:.1f} ms wall, {st['gpu_ms']:.1f} ms gpu, "
                  f"{sc.width * sc.height * spp / st['gpu_ms'] * 1e-3:.2f} Msamples/s, mean {img.mean():.4f}, "
                  f"extend {st['extend_rays'] / st['paths']:.2f}/sample shadow {st['shadow_rays'] / st['paths']:.2f}/sample")
        ctx.close()


if __name__ == "__main__":
    main()

"""Development aid: do two independent wavefronts on two streams overlap usefully?  Two contexts,
two host threads, each renders Cornell 1024x1024 @ spp; aggregate Mpaths/s vs one context."""
import sys, threading, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pif = int(sys.argv[3]) if len(sys.argv) > 3 else (4 << 20) // nctx
sc = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene").with_size(1024, 1024)
ctxs = []
for k in range(nctx):
    c = api.Context(0); c.upload(sc); c.configure(pif, False); c.render_path(4, seed=1); ctxs.append(c)
bar = threading.Barrier(nctx + 1)
def work(c, k):
    bar.wait()
    c.render_path(spp, seed=10 + k)
    bar.wait()
ths = [threading.Thread(target=work, args=(c, k)) for k, c in enumerate(ctxs)]
for t in ths: t.start()
bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
for t in ths: t.join()
import os
print(f"[{nctx} ctx, pif {pif}, GRID_DIV={os.environ.get('TUTU_GRID_DIV','-')}] aggregate {nctx * 1024 * 1024 * spp / dt * 1e-6:.1f} Mpaths/s; per-ctx gpu_ms {[round(c.stats()['gpu_ms'],1) for c in ctxs]}")

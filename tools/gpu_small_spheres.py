"""Dev: mismatch fraction vs the oracle of the small scene with spheres under the different traversal paths."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from tuturenderer_b200 import api
from oracle import oracle_py as oracle
from test_gpu_render import _cornell_with_spheres
cornell = api.Scene.load(ROOT / "tests/golden/cornell_256.tscene")
sc = _cornell_with_spheres(api, cornell).with_size(72, 72)
o, cnt = oracle.OracleScene(sc).render_path(12, seed=31, counters=True)
ctx = api.Context(0)
ctx.upload(sc)
for name in ("wavefront", "resident"):
    ctx.pipeline(name)
    g = ctx.render_path(12, seed=31)
    d = np.abs(g - o)
    st = ctx.stats()
    print(os.environ.get("TUTU_NO_SMALL"), name, "mismatch", (d > 1e-3 * (1 + np.abs(o))).any(-1).mean(), "median", np.median(d), "rays", st["extend_rays"], cnt[0], st["shadow_rays"], cnt[1],
          "mean ratio", g.mean() / o.mean())
    if os.environ.get("TUTU_NO_SMALL"): break

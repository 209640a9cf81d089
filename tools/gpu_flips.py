"""Dev: fraction of pixels that differ from the same-stream CPU restatement (tests' criterion), per scene."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from tuturenderer_b200 import api
from oracle import oracle_py as oracle
g = ROOT / "tests/golden"
ctx = api.Context(0)
ctx.pipeline("wavefront")
for name, size, spp in (("cornell_256", 128, 16), ("mixed", 96, 16), ("glass_c4", 64, 16)):
    sc = api.Scene.load(g / f"{name}.tscene").with_size(size, size)
    ctx.upload(sc)
    a = ctx.render_path(spp, seed=21)
    o = oracle.OracleScene(sc).render_path(spp, seed=21)
    d = np.abs(a - o)
    print(name, f"pixels off by > 1e-3 rel: {(d > 1e-3 * (1 + np.abs(o))).any(-1).mean():.4%}, > 1e-5 rel: {(d > 1e-5 * (1 + np.abs(o))).any(-1).mean():.4%}, median abs diff {np.median(d):.2e}", flush=True)

"""Short fixed workload for ncu: the configs[3] stand-in scene, 1024x1024 @ 8 spp."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api
sc = api.Scene.load(ROOT / "tests/golden/glass_c4.tscene").with_size(1024, 1024)
ctx = api.Context(0)
ctx.upload(sc)
ctx.configure(0, False, 1)
img = ctx.render_path(8, seed=5)
print("glass mean", float(img.mean()), ctx.stats())

"""Glass / texture scene throughput + stage times for the library picked by TUTU_LIB (knobs from the environment)."""
import os, sys, json
sys.path.insert(0, '/root/repo')
from tuturenderer_b200 import api
sc = api.Scene.load('/root/repo/tests/golden/glass_c4.tscene').with_size(1024, 1024)
ctx = api.Context(0); ctx.upload(sc)
for k in range(2):
    ctx.render_path(32, seed=k)
img = ctx.render_path(128, seed=9)
out = {k[5:].lower(): v for k, v in os.environ.items() if k.startswith('TUTU_') and k != 'TUTU_LIB'}
out['glass'] = round(1024 * 1024 * 128 / ctx.stats()['gpu_ms'] * 1e-3, 1); out['mean'] = round(float(img.mean()), 6)
ctx.configure(0, True, 1)
ctx.render_path(64, seed=3)
st = ctx.stats()
out['stages_64spp_1lane'] = [round(st[k], 2) for k in ('extend_ms', 'shade_ms', 'shadow_ms', 'other_ms')]
print(json.dumps(out), flush=True)

"""Throughput of the queue tracers for the library picked by TUTU_LIB (experiment builds; knobs from the environment):
the glass / texture scene (path tracer, stage times of a profiled one-lane pass) and the Veach room BDPT.  One line."""
import os, sys, json
sys.path.insert(0, '/root/repo')
from tuturenderer_b200 import api
G = '/root/repo/tests/golden/'
out = {k[5:].lower(): v for k, v in os.environ.items() if k.startswith('TUTU_') and k != 'TUTU_LIB'}
sc = api.Scene.load(G + 'glass_c4.tscene').with_size(1024, 1024)
ctx = api.Context(0); ctx.upload(sc)
for k in range(2):
    ctx.render_path(32, seed=k)
img = ctx.render_path(128, seed=9)
st = ctx.stats()
out['glass'] = round(1024 * 1024 * 128 / st['gpu_ms'] * 1e-3, 1)
out['glass_mean'] = round(float(img.mean()), 6)
out['glass_rays'] = [st.get('extend_rays'), st.get('shadow_rays')]
ctx.configure(0, True, 1)
ctx.render_path(64, seed=3)
st = ctx.stats()
out['glass_stages_64spp'] = [round(st[k], 2) for k in ('extend_ms', 'shade_ms', 'shadow_ms', 'other_ms')]
ctx.close()
sc = api.Scene.load(G + 'veach_80x60.tscene').with_size(800, 600)
ctx = api.Context(0); ctx.upload(sc)
for k in range(2):
    ctx.render_bdpt(16, seed=k)
im = ctx.render_bdpt(128, seed=9)
out['bdpt'] = round(800 * 600 * 128 / ctx.stats()['gpu_ms'] * 1e-3, 1)
out['bdpt_mean'] = round(float(im.mean()), 6)
ctx.close()
print(json.dumps(out), flush=True)

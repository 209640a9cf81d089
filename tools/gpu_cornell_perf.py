"""Cornell 1024^2 throughput (the headline step) for the library picked by TUTU_LIB: 3 renders of `spp` samples."""
import os, sys, json
sys.path.insert(0, '/root/repo')
from tuturenderer_b200 import api
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sc = api.Scene.load('/root/repo/tests/golden/cornell_256.tscene').with_size(1024, 1024)
ctx = api.Context(0); ctx.upload(sc)
for k in range(2):
    ctx.render_path(64, seed=k)
res = []
for k in range(3):
    img = ctx.render_path(spp, seed=10 + k)
    st = ctx.stats()
    res.append(round(1024 * 1024 * spp / st['gpu_ms'] * 1e-3, 1))
ctx.configure(0, True, 1)
ctx.render_path(64, seed=3)
st = ctx.stats()
print(json.dumps({'lib': os.environ.get('TUTU_LIB', 'default').split('libtutu_b200')[-1], 'mpaths_s': res, 'mean': float(img.mean()),
                  'stages_ms_64spp_1lane': {k: round(st[k + '_ms'], 2) for k in ('extend', 'shade', 'shadow', 'other')}}), flush=True)

"""Cornell 1024^2: throughput with 1, 2 and 3 wavefront lanes (no stage profiling) for the library picked by TUTU_LIB."""
import os, sys, json
sys.path.insert(0, '/root/repo')
from tuturenderer_b200 import api
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sc = api.Scene.load('/root/repo/tests/golden/cornell_256.tscene').with_size(1024, 1024)
ctx = api.Context(0); ctx.upload(sc)
out = {'lib': os.environ.get('TUTU_LIB', 'default').split('libtutu_b200')[-1], 'block': os.environ.get('TUTU_SHADE_BLOCK_RT', '')}
for lanes in (1, 2, 3):
    ctx.configure(0, False, lanes)
    ctx.render_path(64, seed=1)
    res = []
    for k in range(2):
        ctx.render_path(spp, seed=10 + k)
        res.append(round(1024 * 1024 * spp / ctx.stats()['gpu_ms'] * 1e-3, 1))
    out['lanes%d' % lanes] = res
print(json.dumps(out), flush=True)

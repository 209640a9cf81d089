"""BDPT throughput on the tree scenes (Veach room, glass / texture scene) for the library picked by TUTU_LIB."""
import os, sys, json
sys.path.insert(0, '/root/repo')
from tuturenderer_b200 import api
G = '/root/repo/tests/golden/'
out = {k[5:].lower(): v for k, v in os.environ.items() if k.startswith('TUTU_') and k != 'TUTU_LIB'}
for name, size in (('veach_80x60', (800, 600)), ('glass_c4', (768, 768))):
    sc = api.Scene.load(G + name + '.tscene').with_size(*size)
    ctx = api.Context(0); ctx.upload(sc)
    for k in range(2):
        ctx.render_bdpt(8, seed=k)
    im = ctx.render_bdpt(64, seed=9)
    st = ctx.stats()
    out[name] = round(size[0] * size[1] * 64 / st['gpu_ms'] * 1e-3, 2); out[name + '_mean'] = round(float(im.mean()), 6)
    out[name + '_measured'] = ctx.bdpt_queue_tracer_measured()
    ctx.render_bdpt(64, seed=10)
    out[name + '_second_render'] = round(size[0] * size[1] * 64 / ctx.stats()['gpu_ms'] * 1e-3, 2)
    ctx.close()
print(json.dumps(out), flush=True)

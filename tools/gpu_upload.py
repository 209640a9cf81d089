"""Upload time breakdown of the 999 698-triangle scene for both traversal-tree builders (5 uploads each)."""
import sys, json
sys.path.insert(0, '/root/repo')
from tuturenderer_b200 import api
prims = api.synth_heightfield(707)
sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
ctx = api.Context(0)
for builder in ('host_sah', 'device_lbvh'):
    ctx.builder(builder)
    for rep in range(5):
        ctx.upload(sc)
        print(builder, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in ctx.upload_stats().items()}, flush=True)

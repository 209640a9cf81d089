"""Mixed-material scene against the reference's 2048-spp mean: worst 8x8 block per seed and the pixels behind it
(the deterministic specular-chain pixels masked by name in tests/test_gpu_render.py)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from tuturenderer_b200 import api
g = '/root/repo/tests/golden/'
ref = np.fromfile(g + 'mixed_96_ref_mean_2048.f32', np.float32).reshape(96, 96, 3)
sc = api.Scene.load(g + 'mixed.tscene')
ctx = api.Context(0); ctx.upload(sc)
b = lambda a: a.reshape(12, 8, 12, 8, 3).mean((1, 3))
for seed in (1, 2, 3):
    img = ctx.render_path(8192, seed=seed)
    rel = np.abs(b(img) - b(ref)) / (b(ref) + 0.02)
    k = np.unravel_index(rel.argmax(), rel.shape)
    print('seed', seed, 'max', rel.max(), 'at block', k, 'mean', rel.mean(), 'img', b(img)[k], 'ref', b(ref)[k])
    d = np.abs(img - ref).max(-1)
    for idx in np.argsort(d.ravel())[::-1][:8]:
        y, x = divmod(int(idx), 96)
        print('   pixel', (y, x), 'gpu', img[y, x], 'ref', ref[y, x], 'absdiff', d[y, x])
    by, bx = k[0], k[1]
    blk = d[by * 8:by * 8 + 8, bx * 8:bx * 8 + 8]
    print('   worst block abs diffs (rows):')
    for row in blk:
        print('     ', ' '.join('%.3f' % v for v in row))

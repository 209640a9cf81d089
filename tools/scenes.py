"""Procedural scenes for tests, goldens and benches (our own geometry, no reference assets).

mixed_scene(): a Cornell-sized room that exercises every primitive / material / texture path of
the hot path (BASELINE.json configs[3] stand-in: the reference ships no bunny and no texture
files): Lambertian walls, a two-triangle ceiling light and a small sphere light,
a PERFECT_REFLECTIVE sphere, a PERFECT_REFRACTIVE sphere, a MICROFACET_T smooth-shaded blob, a
MICROFACET_R box with albedo/normal/roughness/metallic maps, a textured Lambertian sphere and an
UNLIT quad.
"""
from __future__ import annotations

import numpy as np

from tuturenderer_b200 import api


def _tri(v, n, uv, material, tex=None):
    p = np.zeros(1, api.PRIM_DTYPE)
    p["type"] = api.PRIM_TRIANGLE
    p["v"] = np.asarray(v, np.float32).reshape(9)
    p["n"] = np.asarray(n, np.float32).reshape(9)
    p["uv"] = np.asarray(uv, np.float32).reshape(6)
    p["material"] = material
    p["tex_active"] = 0
    p["tex_diffuse"] = p["tex_normal"] = p["tex_roughness"] = p["tex_metallic"] = -1
    if tex is not None:
        p["tex_active"] = 1
        p["tex_diffuse"], p["tex_normal"], p["tex_roughness"], p["tex_metallic"] = tex
    return p


def _flat_normal(a, b, c):
    n = np.cross(np.asarray(b, np.float64) - a, np.asarray(c, np.float64) - a)
    return (n / np.linalg.norm(n)).astype(np.float32)


def quad(p0, p1, p2, p3, material, tex=None, flip=False):
    """Two triangles (p0,p1,p2), (p0,p2,p3) with flat normals and uv (0,0),(1,0),(1,1),(0,1)."""
    pts = [np.asarray(p, np.float32) for p in (p0, p1, p2, p3)]
    n = _flat_normal(pts[0], pts[1], pts[2])
    if flip:
        n = -n
    uv = [(0, 0), (1, 0), (1, 1), (0, 1)]
    out = []
    for idx in ((0, 1, 2), (0, 2, 3)):
        out.append(_tri([pts[i] for i in idx], [n] * 3, [uv[i] for i in idx], material, tex))
    return out


def box(lo, hi, material, tex=None):
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    f = []
    f += quad((x0, y0, z0), (x0, y1, z0), (x1, y1, z0), (x1, y0, z0), material, tex)  # front  (-z)
    f += quad((x1, y0, z1), (x1, y1, z1), (x0, y1, z1), (x0, y0, z1), material, tex)  # back   (+z)
    f += quad((x0, y0, z1), (x0, y1, z1), (x0, y1, z0), (x0, y0, z0), material, tex)  # left   (-x)
    f += quad((x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1), material, tex)  # right  (+x)
    f += quad((x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0), material, tex)  # top    (+y)
    f += quad((x0, y0, z1), (x0, y0, z0), (x1, y0, z0), (x1, y0, z1), material, tex)  # bottom (-y)
    return f


def sphere(center, radius, material, tex=None):
    p = np.zeros(1, api.PRIM_DTYPE)
    p["type"] = api.PRIM_SPHERE
    p["v"][0, 0:3] = center
    p["v"][0, 3] = radius
    p["material"] = material
    p["tex_active"] = 0
    p["tex_diffuse"] = p["tex_normal"] = p["tex_roughness"] = p["tex_metallic"] = -1
    if tex is not None:
        p["tex_active"] = 1
        p["tex_diffuse"], p["tex_normal"], p["tex_roughness"], p["tex_metallic"] = tex
    return [p]


def blob(center, radius, material, subdiv=2):
    """Subdivided octahedron pushed to a bumpy sphere, smooth (per-vertex) normals."""
    v = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    faces = [(0, 2, 4), (2, 1, 4), (1, 3, 4), (3, 0, 4), (2, 0, 5), (1, 2, 5), (3, 1, 5), (0, 3, 5)]
    tris = [[np.asarray(v[i], np.float64) for i in f] for f in faces]
    for _ in range(subdiv):
        nxt = []
        for a, b, c in tris:
            ab, bc, ca = (a + b) / 2, (b + c) / 2, (c + a) / 2
            nxt += [[a, ab, ca], [ab, b, bc], [ca, bc, c], [ab, bc, ca]]
        tris = nxt
    out = []
    c = np.asarray(center, np.float64)
    for t in tris:
        pts, nrm = [], []
        for p in t:
            d = p / np.linalg.norm(p)
            r = radius * (1.0 + 0.12 * np.sin(5 * d[0]) * np.cos(4 * d[1] + 1.0))
            pts.append((c + r * d).astype(np.float32))
            nrm.append(d.astype(np.float32))
        out.append(_tri(pts, nrm, [(0, 0), (1, 0), (0, 1)], material))
    return out


def _textures(res=64):
    y, x = np.mgrid[0:res, 0:res].astype(np.float32) / res
    checker = (((x * 8).astype(int) + (y * 8).astype(int)) % 2).astype(np.float32)
    albedo = np.stack([0.25 + 0.6 * checker, 0.3 + 0.3 * (1 - checker), 0.2 + 0.5 * x], -1)
    nx = 0.35 * np.sin(2 * np.pi * 6 * x)
    ny = 0.35 * np.cos(2 * np.pi * 5 * y)
    nz = np.sqrt(np.clip(1 - nx * nx - ny * ny, 0.05, 1))
    normal = np.stack([nx, ny, nz], -1)  # already in [-1,1] (PPMGenerator.hpp:714-720 rescales at load)
    rough = np.repeat((0.25 + 0.5 * x)[..., None], 3, -1)
    metal = np.repeat(checker[..., None] * 0.9, 3, -1)
    return [[albedo.astype(np.float32)], [normal.astype(np.float32)], [rough.astype(np.float32)],
            [metal.astype(np.float32)]]


def mixed_scene(width=96, height=96) -> api.Scene:
    M = api.default_material
    mats = np.concatenate([
        M(diffuse=(0.725, 0.71, 0.68)),                                          # 0 white
        M(diffuse=(0.63, 0.065, 0.05)),                                          # 1 red
        M(diffuse=(0.14, 0.45, 0.091)),                                          # 2 green
        M(diffuse=(0.725, 0.71, 0.68), emission=(47.8348007, 38.5663986, 31.0807991)),  # 3 area light
        M(diffuse=(0.9, 0.9, 0.9), emission=(30.0, 30.0, 60.0)),                  # 4 sphere light
        M(type=api.MAT_PERFECT_REFLECTIVE, diffuse=(0.9, 0.9, 0.9)),             # 5 mirror
        M(type=api.MAT_PERFECT_REFRACTIVE, eta=1.5),                             # 6 glass
        M(type=api.MAT_MICROFACET_T, eta=1.5, roughness=0.35),                   # 7 rough glass
        M(type=api.MAT_MICROFACET_R, diffuse=(0.8, 0.6, 0.2), roughness=0.4, metallic=0.5),  # 8 textured GGX
        M(type=api.MAT_UNLIT, diffuse=(0.1, 0.4, 0.9)),                          # 9 unlit
        M(type=api.MAT_MICROFACET_R, diffuse=(0.95, 0.64, 0.54), roughness=0.25, metallic=1.0),  # 10 copper
    ])
    P = []
    S = 556.0
    P += quad((0, 0, 0), (S, 0, 0), (S, 0, S), (0, 0, S), 0, flip=True)      # floor (normal +y)
    P += quad((0, S, 0), (0, S, S), (S, S, S), (S, S, 0), 0, flip=True)      # ceiling (normal -y)
    P += quad((0, 0, S), (S, 0, S), (S, S, S), (0, S, S), 0, flip=True)      # back wall
    P += quad((0, 0, 0), (0, 0, S), (0, S, S), (0, S, 0), 1, flip=True)      # one side
    P += quad((S, 0, 0), (S, S, 0), (S, S, S), (S, 0, S), 2, flip=True)      # other side
    P += quad((213, S - 0.2, 227), (343, S - 0.2, 227), (343, S - 0.2, 332), (213, S - 0.2, 332), 3)  # light, faces down
    P += sphere((90, 430, 300), 22, 4)                                        # sphere light
    P += sphere((140, 90, 330), 90, 5)                                        # mirror sphere
    P += sphere((400, 80, 150), 80, 6)                                        # glass sphere
    P += blob((300, 260, 260), 70, 7, subdiv=2)                               # rough-glass blob (128 tris)
    P += box((330, 0, 330), (470, 200, 470), 8, tex=(0, 0, 0, 0))             # textured GGX box
    P += sphere((90, 60, 120), 60, 0, tex=(0, 0, -1, -1))                     # textured Lambertian sphere
    P += quad((200, 300, S - 1), (300, 300, S - 1), (300, 380, S - 1), (200, 380, S - 1), 9)  # unlit panel
    P += box((230, 0, 60), (310, 60, 140), 10)                                # copper block
    prims = np.concatenate(P)
    return api.Scene(prims=prims, materials=mats, eye=(278, 273, -800), viewdir=(0, 0, 1), updir=(0, 1, 0),
                     hfov_deg=40, width=width, height=height, bkgcolor=(0.05, 0.06, 0.08), eta=1.0,
                     textures=_textures())


def glass_scene(width=96, height=96, golden_dir=None) -> api.Scene:
    """BASELINE.json configs[3] stand-in ("glass bunny scene with GGX microfacet transmission +
    textured materials"; the reference ships neither a bunny nor texture files):
    the Cornell shell, light and tall box of src/main_cornellBox.cpp, the 1214-triangle smooth-shaded
    glass object of model/veach_bdpt/veach_glass.obj scaled x300 and moved onto the floor
    (PPMGenerator::scaleObj / transObj arithmetic: v * s, then v + t in fp32) as MICROFACET_T
    (eta 1.5, roughness 0.2), and a MICROFACET_R box carrying albedo / normal / roughness / metallic
    maps.  Geometry comes from the committed fixtures (which the reference's own loader produced)."""
    from pathlib import Path
    g = Path(golden_dir) if golden_dir else Path(__file__).resolve().parent.parent / "tests" / "golden"
    cornell = api.Scene.load(g / "cornell_256.tscene")
    veach = api.Scene.load(g / "veach_80x60.tscene")
    M = api.default_material
    mats = np.concatenate([
        cornell.materials,                                                        # 0 white 1 light 2 green 3 red
        M(type=api.MAT_MICROFACET_T, eta=1.5, roughness=0.2),                     # 4 rough glass
        M(type=api.MAT_MICROFACET_R, diffuse=(0.8, 0.6, 0.2), roughness=0.4, metallic=0.5),  # 5 textured GGX
    ])
    shell = cornell.prims[:22].copy()  # floor, ceiling, back wall, light, side walls, tall box
    glass = veach.prims[veach.prims["material"] == 4].copy()
    v = glass["v"].reshape(-1, 3, 3)
    lo, hi = v.min((0, 1)), v.max((0, 1))
    scale = np.float32(300.0)
    offset = np.array([185.0, 0.5, 169.0], np.float32) - np.array([(lo[0] + hi[0]) / 2, lo[1], (lo[2] + hi[2]) / 2], np.float32) * scale
    glass["v"] = (v * scale + offset).astype(np.float32).reshape(-1, 9)
    glass["material"] = 4
    tbox = np.concatenate(box((60, 0, 330), (180, 120, 450), 5, tex=(0, 0, 0, 0)))
    prims = np.concatenate([shell, glass, tbox])
    return api.Scene(prims=prims, materials=mats, eye=(278, 273, -800), viewdir=(0, 0, 1), updir=(0, 1, 0),
                     hfov_deg=40, width=width, height=height, bkgcolor=(0, 0, 0), eta=1.0, textures=_textures())

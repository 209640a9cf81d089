"""Generates tests/golden/* by running the UNMODIFIED reference (oracle/_ref/ref_harness, built by
oracle/Makefile from /root/reference) in this container.  The fixtures travel to the GPU box,
/root/reference does not.  Re-run:  python tests/tools/make_golden.py [--renders]

Fixtures:
  cornell_256.tscene         scene of src/main_cornellBox.cpp via objl::Loader + loadObj, with the
                             reference-built BVH topology (pre-order)
  cornell_rays.f32           deterministic ray batch (64x64 primary rays + rays from inside the box)
  cornell_closest.bin        getIntersection per ray {prim,t,u,v}
  cornell_any.bin            hasIntersection per ray
  hf24.tscene / hf24_*       24x24 height-field (1152 triangles) + both synthetic ray kinds
  mixed.tscene / mixed_*     spheres + triangles, every material type, 4 texture channels
  *_ref_mean_*.f32, stats.json   (with --renders) high-spp reference renders and noise statistics
  veach_80x60.tscene, *_bdpt_ref_mean_*.f32   (with --bdpt) src/main_veach_bdpt.cpp's scene and
                             high-spp means of the reference's BDPT integrator
"""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tuturenderer_b200 import api
from oracle import oracle_py as O
G = ROOT / "tests" / "golden"


def cornell_rays(scene):
    s64 = scene.with_size(64, 64)
    prim = O.OracleScene(s64).primary_rays()
    rng = np.random.default_rng(2024)
    n = 12000
    extra = np.zeros((n, 8), np.float32)
    extra[:, 0:3] = rng.uniform([0, 0, 0], [556, 548, 559], (n, 3))
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    extra[:, 4:7] = d
    extra[:, 7] = rng.uniform(5, 700, n)
    # axis-parallel rays (zero direction components -> inf/NaN slabs) from points on box planes
    ax = np.zeros((600, 8), np.float32)
    ax[:, 0:3] = rng.choice([0.0, 130.0, 278.0, 548.8, 556.0, 559.2], (600, 3))
    k = rng.integers(0, 3, 600)
    ax[np.arange(600), 4 + k] = rng.choice([-1.0, 1.0], 600)
    ax[:, 7] = 400.0
    prim[:, 7] = 900.0
    return np.concatenate([prim, extra, ax]).astype(np.float32)


def main():
    G.mkdir(parents=True, exist_ok=True)
    O.build(ref=True)
    assert O.ref_available(), "oracle/_ref/ref_harness missing (needs /root/reference)"
    O.ref_dump_cornell(256, 256, G / "cornell_256.tscene")
    sc = api.Scene.load(G / "cornell_256.tscene")
    rays = cornell_rays(sc)
    rays.tofile(G / "cornell_rays.f32")
    O.ref_trace(sc, rays, "closest")[0].tofile(G / "cornell_closest.bin")
    O.ref_trace(sc, rays, "any")[0].tofile(G / "cornell_any.bin")
    print("cornell:", len(sc.prims), "prims", len(rays), "rays")

    # height-field
    prims = api.synth_heightfield(24, 12345)
    hf = api.Scene(prims=prims, materials=api.default_material(), width=32, height=32,
                   eye=(0.5, 1.5, 0.5), viewdir=(0, -1, 0), updir=(0, 0, 1))
    hf = O.ref_export_bvh(hf, G / "hf24.tscene")
    for kind in (0, 1):
        r = api.synth_rays(kind, 6000, 12345)
        r.tofile(G / f"hf24_rays{kind}.f32")
        O.ref_trace(hf, r, "closest")[0].tofile(G / f"hf24_closest{kind}.bin")
        O.ref_trace(hf, r, "any")[0].tofile(G / f"hf24_any{kind}.bin")
    print("hf24:", len(hf.prims), "prims")

    from tools.scenes import mixed_scene
    mx = O.ref_export_bvh(mixed_scene(96, 96), G / "mixed.tscene")
    rng = np.random.default_rng(7)
    n = 8000
    r = np.zeros((n, 8), np.float32)
    r[:, 0:3] = rng.uniform([20, 20, 20], [530, 530, 530], (n, 3))
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r[:, 4:7] = d.astype(np.float32)
    # Sphere::intersect assumes unit directions; renormalise in float like the reference would
    r[:, 7] = rng.uniform(5, 700, n)
    r = np.concatenate([O.OracleScene(mx).primary_rays(), r]).astype(np.float32)
    r[:, 7] = np.where(r[:, 7] > 1e30, 900.0, r[:, 7])
    r.tofile(G / "mixed_rays.f32")
    O.ref_trace(mx, r, "closest")[0].tofile(G / "mixed_closest.bin")
    O.ref_trace(mx, r, "any")[0].tofile(G / "mixed_any.bin")
    print("mixed:", len(mx.prims), "prims")

    if "--renders" in sys.argv:
        stats = {}
        c128 = sc.with_size(128, 128)
        runs = [O.ref_render(c128, 2048)[0] for _ in range(2)]
        mean = (runs[0] + runs[1]) * 0.5
        mean.astype(np.float32).tofile(G / "cornell_128_ref_mean_4096.f32")
        def rmse(a, b): return float(np.sqrt(((a - b) ** 2).mean()))
        def relmse(a, b): return float((((a - b) ** 2) / (b ** 2 + 1e-2)).mean())
        stats["cornell_128"] = {"ref_spp_total": 4096, "image_mean": float(mean.mean()),
                                "channel_means": [float(x) for x in mean.mean((0, 1))],
                                "run_to_run_rmse_2048": rmse(runs[0], runs[1])}
        for spp in (16, 64):
            imgs = [O.ref_render(c128, spp)[0] for _ in range(4)]
            stats["cornell_128"][f"rmse_{spp}"] = float(np.mean([rmse(i, mean) for i in imgs]))
            stats["cornell_128"][f"relmse_{spp}"] = float(np.mean([relmse(i, mean) for i in imgs]))
        m96 = [O.ref_render(mx, 1024)[0] for _ in range(2)]
        mm = (m96[0] + m96[1]) * 0.5
        mm.astype(np.float32).tofile(G / "mixed_96_ref_mean_2048.f32")
        stats["mixed_96"] = {"ref_spp_total": 2048, "image_mean": float(np.nanmean(mm)),
                             "run_to_run_rmse_1024": rmse(m96[0], m96[1])}
        (G / "stats.json").write_text(json.dumps(stats, indent=1))
        print(json.dumps(stats, indent=1))


def bdpt_goldens():
    """BDPT fixtures (BASELINE.json configs[4]): the Veach-room scene of src/main_veach_bdpt.cpp with
    the reference-built BVH, and high-spp means of the reference's BDPT for Cornell and Veach."""
    O.build(ref=True)
    assert O.ref_available()
    O.ref_dump_veach(80, 60, G / "veach_80x60.tscene")
    veach = api.Scene.load(G / "veach_80x60.tscene")
    cornell = api.Scene.load(G / "cornell_256.tscene").with_size(64, 64)
    stats = json.loads((G / "stats.json").read_text())
    def rmse(a, b): return float(np.sqrt(((a - b) ** 2).mean()))
    for name, sc, spp in (("cornell_64_bdpt", cornell, 2048), ("veach_80x60_bdpt", veach, 1024)):
        runs = [O.ref_render(sc, spp, mode="bdpt-rows")[0] for _ in range(2)]
        mean = (runs[0] + runs[1]) * 0.5
        mean.astype(np.float32).tofile(G / f"{name}_ref_mean_{2 * spp}.f32")
        low = [O.ref_render(sc, 16, mode="bdpt-rows")[0] for _ in range(4)]
        stats[name] = {"ref_spp_total": 2 * spp, "image_mean": float(mean.mean()),
                       "channel_means": [float(x) for x in mean.mean((0, 1))],
                       f"run_to_run_rmse_{spp}": rmse(runs[0], runs[1]),
                       "rmse_16": float(np.mean([rmse(i, mean) for i in low]))}
        print(name, stats[name])
    (G / "stats.json").write_text(json.dumps(stats, indent=1))


def c4_goldens():
    """configs[3] stand-in (tools/scenes.py: glass_scene): reference-built BVH + a high-spp mean of the
    reference's PathTracing."""
    from tools.scenes import glass_scene
    O.build(ref=True)
    stats = json.loads((G / "stats.json").read_text())
    def rmse(a, b): return float(np.sqrt(((a - b) ** 2).mean()))
    if "--bdpt-only" in sys.argv:
        sc = api.Scene.load(G / "glass_c4.tscene")
        stats_pt = stats["glass_c4_96"]
    else:
        sc = O.ref_export_bvh(glass_scene(96, 96), G / "glass_c4.tscene")
        runs = [O.ref_render(sc, 1024)[0] for _ in range(2)]
        mean = (runs[0] + runs[1]) * 0.5
        mean.astype(np.float32).tofile(G / "glass_c4_96_ref_mean_2048.f32")
    stats["glass_c4_96"] = stats_pt if "--bdpt-only" in sys.argv else {"ref_spp_total": 2048, "image_mean": float(np.nanmean(mean)),
                            "channel_means": [float(x) for x in np.nanmean(mean, (0, 1))],
                            "run_to_run_rmse_1024": rmse(runs[0], runs[1]), "nan_pixels": int(np.isnan(mean).any(-1).sum())}
    print(stats["glass_c4_96"])
    # the same scene through the reference's BDPT (64x64): MICROFACET_T / MICROFACET_R / textures in buildEyePath,
    # buildLightPath (adjoint BSDF) and MISweight
    sc64 = sc.with_size(64, 64)
    runs = [O.ref_render(sc64, 512, mode="bdpt-rows")[0] for _ in range(2)]
    mean = (runs[0] + runs[1]) * 0.5
    mean.astype(np.float32).tofile(G / "glass_c4_64_bdpt_ref_mean_1024.f32")
    stats["glass_c4_64_bdpt"] = {"ref_spp_total": 1024, "image_mean": float(np.nanmean(mean)),
                                 "channel_means": [float(x) for x in np.nanmean(mean, (0, 1))],
                                 "run_to_run_rmse_512": rmse(runs[0], runs[1]),
                                 "median_abs_run_to_run_512": float(np.median(np.abs(runs[0] - runs[1])))}
    print(stats["glass_c4_64_bdpt"])
    (G / "stats.json").write_text(json.dumps(stats, indent=1))


def ppm_golden():
    """Output stage: a 64x48 float image covering [0,1] densely plus negatives, > 1, NaN and inf,
    written by the reference's own PPMGenerator::generate."""
    O.build(ref=True)
    rng = np.random.default_rng(5)
    img = rng.uniform(-0.1, 1.2, (48, 64, 3)).astype(np.float32)
    img[0, :, 0] = np.linspace(0, 1, 64, dtype=np.float32)
    img[1, :8, 1] = [np.nan, np.inf, -np.inf, 0.0, 1.0, -0.0, 1e-30, 0.999999]
    img[2] = (np.arange(64 * 3, dtype=np.float32).reshape(64, 3) / 255.0) ** np.float32(1 / 0.78)  # near level boundaries
    img.tofile(G / "ppm_in_64x48.f32")
    O.ref_ppm(img, G / "ppm_ref_64x48.ppm")
    print("ppm golden written")


def c256_goldens():
    """SURVEY.md §8(d) gates at their stated size: Cornell 256x256.  R* = mean of four independent 1024-spp reference
    runs; the reference's own noise level at N = 16 and N = 1024 spp against it (N = 1024 leave-one-out: each run
    against the mean of the other three), its run-to-run RMSE at 1024 spp, and the deterministic pixel masks."""
    O.build(ref=True)
    stats = json.loads((G / "stats.json").read_text())
    def rmse(a, b): return float(np.sqrt(((a - b) ** 2).mean()))
    def relmse(a, b): return float((((a - b) ** 2) / (b ** 2 + 1e-2)).mean())
    file = G / "cornell_256.tscene"
    runs = [O.ref_render_file(file, 1024, width=256, height=256)[0].astype(np.float64) for _ in range(4)]
    mean4 = sum(runs) / 4
    mean4.astype(np.float32).tofile(G / "cornell_256_ref_mean_4096.f32")
    loo = [(sum(runs) - r) / 3 for r in runs]  # mean of the other three
    # the first three runs' mean is stored too: a GPU 1024-spp render against it has the same expected RMSE as the
    # reference's leave-one-out figure (sigma^2 / 1024 + sigma^2 / 3072)
    (sum(runs[:3]) / 3).astype(np.float32).tofile(G / "cornell_256_ref_mean_3072.f32")
    imgs16 = [O.ref_render_file(file, 16, width=256, height=256)[0] for _ in range(4)]
    stats["cornell_256"] = {
        "ref_spp_total": 4096, "image_mean": float(mean4.mean()), "channel_means": [float(x) for x in mean4.mean((0, 1))],
        "rmse_16": float(np.mean([rmse(i, mean4) for i in imgs16])), "relmse_16": float(np.mean([relmse(i, mean4) for i in imgs16])),
        "rmse_1024_leave_one_out": float(np.mean([rmse(r, m) for r, m in zip(runs, loo)])),
        "relmse_1024_leave_one_out": float(np.mean([relmse(r, m) for r, m in zip(runs, loo)])),
        "run_to_run_rmse_1024": float(np.mean([rmse(runs[0], runs[1]), rmse(runs[2], runs[3])])),
        "background_pixels": int((mean4 == 0).all(-1).sum()),
        # pixels whose primary ray hits the light: every sample is the emission (summed in fp32: equal to 1e-4, not bitwise)
        "emission_pixels": int(np.isclose(mean4, np.float32([47.8348007, 38.5663986, 31.0807991]), rtol=1e-4, atol=0).all(-1).sum())}
    print(stats["cornell_256"])
    (G / "stats.json").write_text(json.dumps(stats, indent=1))


def post_inputs():
    """Postprocessor fixtures' inputs: (a) 40 rows x 48 columns with emissive patches (|rgb| > 3) in the interior, in
    the corners (the u == 0 / v == 0 wrap-around of Texture::getRGBat) and values around the tone map's
    cancellation range; (b) an 8 x 1030 strip: frames wider than 1000 pixels reach clamp(0, 0.999, u)."""
    rng = np.random.default_rng(7)
    a = rng.uniform(0, 1.2, (40, 48, 3)).astype(np.float32)
    a[10:14, 20:26] = [10, 8, 6]
    a[0, 0] = [5, 5, 5]
    a[39, 47] = [0, 9, 0]
    a[5, 0] = [4, 0, 0]
    a[30, 40:44] = [1.8, 1.8, 1.7]  # norm just above / below 3
    a[31, 40:44] = [1.7, 1.7, 1.7]
    a[20, :7, 0] = [0, 1e-4, 3e-3, 1e-6, 0.5, 2.9, 1.74]
    b = rng.uniform(0, 1.0, (8, 1030, 3)).astype(np.float32)
    b[3, 1020:1030] = [6, 6, 6]
    b[7, 0:3] = [0, 0, 7]
    return a, b


def post_golden():
    """Output stage, Postprocessor.hpp:29-197 run by the UNMODIFIED reference (ref_harness postprocess)."""
    O.build(ref=True)
    a, b = post_inputs()
    a.tofile(G / "post_in_48x40.f32")
    b.tofile(G / "post_in_1030x8.f32")
    for mode in ("extract", "blur", "bloom", "hdr", "full"):
        out, _ = O.ref_postprocess(a, mode)
        out.tofile(G / f"post_ref_48x40_{mode}.f32")
    for mode in ("blur", "full"):
        out, _ = O.ref_postprocess(b, mode)
        out.tofile(G / f"post_ref_1030x8_{mode}.f32")
    print("postprocess goldens written")


if __name__ == "__main__":
    if "--c256" in sys.argv:
        c256_goldens()
    elif "--post" in sys.argv:
        post_golden()
    elif "--ppm" in sys.argv:
        ppm_golden()
    elif "--c4" in sys.argv:
        c4_goldens()
    elif "--bdpt" in sys.argv:
        bdpt_goldens()
    else:
        main()

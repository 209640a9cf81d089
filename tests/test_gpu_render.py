"""Parity tests proper for the wavefront path tracer (through the C ABI, on a B200).

Two kinds of gate:
* same-stream: the oracle draws the same Philox numbers, so GPU and oracle trace the same paths;
  images agree to float noise except where a 1-ulp difference flips a discrete decision;
* statistical: against the reference's own converged render and noise figures (tests/golden,
  produced by the unmodified reference), SURVEY.md §8d gates 1-3.
Tolerances (north_star: "within a stated per-pixel RMSE / relative-MSE tolerance") are written
next to each assertion."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LIGHT = np.array([47.8348007, 38.5663986, 31.0807991], np.float32)


@pytest.fixture(params=["auto", "wavefront"])
def ctx(api, request):
    """Every test of this module runs with the automatic pipeline choice (small renders of the Cornell scenes
    take the register-resident kernel) and with the wavefront forced."""
    c = api.Context(0)
    c.pipeline(request.param)
    yield c
    c.close()


def _rmse(a, b):
    return float(np.sqrt(((a - b) ** 2).mean()))


def _relmse(a, b):
    return float((((a - b) ** 2) / (b ** 2 + 1e-2)).mean())


# the two larger cases give wf_shade's blocks enough iterations to append through their reserves (wavefront.cuh:
# "queue appends"; a lost or doubled queue entry would show in the ray counts and in the image)
@pytest.mark.parametrize("name,size,spp", [("cornell_256", 64, 8), ("mixed", 48, 16), ("cornell_256", 320, 10), ("mixed", 224, 6)])
def test_same_stream_as_oracle(api, oracle, ctx, golden, name, size, spp):
    sc = api.Scene.load(golden / f"{name}.tscene").with_size(size, size)
    ctx.upload(sc)
    g = ctx.render_path(spp, seed=21)
    o, cnt = oracle.OracleScene(sc).render_path(spp, seed=21, counters=True)
    assert np.isfinite(g).all() == np.isfinite(o).all()
    d = np.abs(g - o)
    # <= 2 % of pixels may differ visibly (discrete flips: libm vs CUDA ulps at branch points)
    assert (d > 1e-3 * (1 + np.abs(o))).any(-1).mean() < 0.02
    assert abs(g.mean() / o.mean() - 1) < 5e-3
    st = ctx.stats()
    assert st["paths"] == size * size * spp
    # the reference recurses out of calcForRefractive BEFORE testing pdf < MIN_DIVISOR
    # (PathTracing.hpp:128-133) and so traces sub-paths it then discards; the GPU tests first.
    # Without refractive materials (Cornell) the ray counts agree to discrete flips.
    assert st["extend_rays"] <= cnt[0] * 1.002
    assert st["extend_rays"] >= cnt[0] * (0.998 if name.startswith("cornell") else 0.97)
    # the CPU traces every NEE shadow ray, the GPU only those whose contribution is not already
    # rejected by the facing tests (PathTracing.hpp:197,202), so it traces at most as many
    assert st["shadow_rays"] <= cnt[1]


def test_cornell_statistics_against_reference(api, ctx, golden, cornell):
    stats = json.loads((golden / "stats.json").read_text())["cornell_128"]
    ref = np.fromfile(golden / "cornell_128_ref_mean_4096.f32", np.float32).reshape(128, 128, 3)
    ctx.upload(cornell.with_size(128, 128))
    # gate 1: noise level at N spp within +-10 % of the reference's own at the same N
    for spp in (16, 64):
        vals_r, vals_m = [], []
        for seed in (1, 2, 3, 4):
            img = ctx.render_path(spp, seed=seed)
            vals_r.append(_rmse(img, ref))
            vals_m.append(_relmse(img, ref))
        assert abs(np.mean(vals_r) / stats[f"rmse_{spp}"] - 1) < 0.10
        assert abs(np.mean(vals_m) / stats[f"relmse_{spp}"] - 1) < 0.10
    # gate 3: deterministic pixel sets.  Where the reference's primary ray misses (exact bkgcolor)
    # or hits the light (exact emission) the GPU pixel must be exactly that at any spp.
    img = ctx.render_path(16, seed=9)
    assert ((img == 0).all(-1) >= (ref == 0).all(-1)).all()
    assert np.array_equal((img == LIGHT).all(-1), (ref == LIGHT).all(-1))
    # gate 2: bias.  16384 spp vs the 4096-spp reference mean: per-channel mean within 0.5 %,
    # per-pixel RMSE below 2x the reference's run-to-run RMSE at 2048 spp
    big = ctx.render_path(16384, seed=5)
    assert np.array_equal((big == 0).all(-1), (ref == 0).all(-1))
    for c in range(3):
        assert abs(big[..., c].mean() / stats["channel_means"][c] - 1) < 0.005
    assert _rmse(big, ref) < 2 * stats["run_to_run_rmse_2048"]


def test_cornell_256_gates_of_the_survey(api, ctx, golden, cornell):
    """SURVEY.md §8(d) correctness gates at the size and sample counts it states: Cornell 256x256, R* = mean of four
    independent 1024-spp REFERENCE runs (tests/tools/make_golden.py --c256).
    Gate 1 (noise level): RMSE and relMSE of GPU renders at N = 16 and at N = 1024 spp against R* within +-10 % of the
    reference's own figures at the same N (N = 1024: the reference's leave-one-out figure, each run against the mean of
    the other three; the GPU render is compared with the stored three-run mean, which has the same expectation).
    Gate 2 (bias): 16 384 GPU spp against R*: channel means within 0.5 %, per-pixel RMSE below 2x the reference's
    run-to-run RMSE at 1024 spp.  Gate 3: background and emission pixel sets identical."""
    st = json.loads((golden / "stats.json").read_text())["cornell_256"]
    ref4 = np.fromfile(golden / "cornell_256_ref_mean_4096.f32", np.float32).reshape(256, 256, 3)
    ref3 = np.fromfile(golden / "cornell_256_ref_mean_3072.f32", np.float32).reshape(256, 256, 3)
    ctx.upload(cornell.with_size(256, 256))
    imgs16 = [ctx.render_path(16, seed=s) for s in (1, 2, 3, 4)]
    assert abs(np.mean([_rmse(i, ref4) for i in imgs16]) / st["rmse_16"] - 1) < 0.10
    assert abs(np.mean([_relmse(i, ref4) for i in imgs16]) / st["relmse_16"] - 1) < 0.10
    imgs1024 = [ctx.render_path(1024, seed=s) for s in (5, 6, 7)]
    assert abs(np.mean([_rmse(i, ref3) for i in imgs1024]) / st["rmse_1024_leave_one_out"] - 1) < 0.10
    assert abs(np.mean([_relmse(i, ref3) for i in imgs1024]) / st["relmse_1024_leave_one_out"] - 1) < 0.10
    big = ctx.render_path(16384, seed=8)
    for c in range(3):
        assert abs(big[..., c].mean() / st["channel_means"][c] - 1) < 0.005
    assert _rmse(big, ref4) < 2 * st["run_to_run_rmse_1024"]
    assert np.array_equal((big == 0).all(-1), (ref4 == 0).all(-1)) and int((big == 0).all(-1).sum()) == st["background_pixels"]
    on_light = lambda a: np.isclose(a, LIGHT, rtol=1e-4, atol=0).all(-1)  # every sample is the emission (fp32 sums of 1024 terms: 1e-4, not bitwise)
    assert np.array_equal(on_light(imgs16[0]), on_light(ref4)) and np.array_equal(on_light(imgs1024[0]), on_light(ref4))
    assert int(on_light(imgs1024[0]).sum()) == st["emission_pixels"] == 380
    # (the 16 384-sample sum of 47.83 loses more than 1e-4 to fp32 rounding; its light pixels are held to 2e-3)
    assert np.array_equal(np.isclose(big, LIGHT, rtol=2e-3, atol=0).all(-1), on_light(ref4))


def test_mixed_scene_statistics_against_reference(api, ctx, golden, mixed):
    """Every material / texture / sphere path (configs[3] stand-in)."""
    stats = json.loads((golden / "stats.json").read_text())["mixed_96"]
    ref = np.fromfile(golden / "mixed_96_ref_mean_2048.f32", np.float32).reshape(96, 96, 3)
    ctx.upload(mixed)
    img = ctx.render_path(8192, seed=2)
    assert np.isfinite(img).all()
    for c in range(3):
        assert abs(img[..., c].mean() / ref[..., c].mean() - 1) < 0.01
    assert _rmse(img, ref) < 2 * stats["run_to_run_rmse_1024"]
    # Block means (8x8) localise a wrong material.  A handful of pixels are excluded first, explicitly: pixels on the
    # silhouettes of the PERFECT_REFLECTIVE / PERFECT_REFRACTIVE spheres where the camera ray (no pixel jitter) starts
    # a (nearly) deterministic specular chain, so a last-place difference in a grazing reflection (FMA contraction /
    # rsqrt in the shading code vs the reference's libm build) moves EVERY sample of the pixel to another surface -
    # e.g. (68, 66): the reference's chain ends in NaN for all 2048 samples (pixel exactly 0, PathTracing.hpp:510)
    # while the GPU's reaches the red wall (0.59).  tools/gpu_mixed_stat.py lists them per seed.  They must be few
    # (<= 8 of 9216), lie on those silhouettes (the named neighbourhoods), and everything else is gated tightly:
    # worst 8x8 block below 10 % (it was a 30 % blanket over all pixels), mean block deviation below 1 %.
    CHAIN_PIXELS = [(68, 66), (70, 69), (42, 39), (43, 38), (40, 46), (40, 47), (57, 47)]
    diff = np.abs(img - ref).max(-1)
    outliers = [tuple(int(v) for v in yx) for yx in np.argwhere(diff > 0.1)]
    assert len(outliers) <= 8, outliers
    for y, x in outliers:
        assert any(abs(y - cy) <= 1 and abs(x - cx) <= 1 for cy, cx in CHAIN_PIXELS), f"unexpected outlier pixel {(y, x)}"
    img_m, ref_m = img.copy(), ref.copy()
    for y, x in outliers:
        img_m[y, x] = ref_m[y, x]
    b = lambda a: a.reshape(12, 8, 12, 8, 3).mean((1, 3))
    rel = np.abs(b(img_m) - b(ref_m)) / (b(ref_m) + 0.02)
    assert rel.max() < 0.10 and rel.mean() < 0.01, (rel.max(), rel.mean())


def test_glass_scene_statistics_against_reference(api, oracle, ctx, golden):
    """BASELINE.json configs[3] stand-in (tools/scenes.py: glass_scene): 1214-triangle MICROFACET_T
    glass object + MICROFACET_R box with all four texture channels inside the Cornell shell, against
    the mean of 2048 reference spp (tests/tools/make_golden.py --c4)."""
    stats = json.loads((golden / "stats.json").read_text())["glass_c4_96"]
    ref = np.fromfile(golden / "glass_c4_96_ref_mean_2048.f32", np.float32).reshape(96, 96, 3)
    sc = api.Scene.load(golden / "glass_c4.tscene")
    ctx.upload(sc)
    img = ctx.render_path(8192, seed=2)
    assert np.isfinite(img).all()
    for c in range(3):  # per-channel mean within 0.5 %
        assert abs(img[..., c].mean() / stats["channel_means"][c] - 1) < 0.005
    assert _rmse(img, ref) < 2 * stats["run_to_run_rmse_1024"]
    b = lambda a: a.reshape(12, 8, 12, 8, 3).mean((1, 3))  # 8x8 block means localise a wrong material
    rel = np.abs(b(img) - b(ref)) / (b(ref) + 0.02)
    assert rel.max() < 0.15 and rel.mean() < 0.01
    # same Philox stream as the oracle on a small frame
    small = sc.with_size(40, 40)
    ctx.upload(small)
    g = ctx.render_path(8, seed=11)
    o = oracle.OracleScene(small).render_path(8, seed=11)
    assert (np.abs(g - o) > 1e-3 * (1 + np.abs(o))).any(-1).mean() < 0.02


def test_sample_ranges_compose(api, ctx, cornell):
    """Samples [0,8) rendered as one call == ranges [0,3) + [3,8) accumulated (multi-GPU split)."""
    import torch
    sc = cornell.with_size(64, 64)
    ctx.upload(sc)
    whole = ctx.render_path(8, seed=4)
    acc = torch.zeros(64 * 64 * 3, dtype=torch.float32, device="cuda")
    out = torch.empty_like(acc)
    torch.cuda.synchronize()  # the library runs on its own stream here (stream = NULL)
    ctx.render_accumulate_device(0, 3, 4, acc.data_ptr())
    ctx.render_accumulate_device(3, 5, 4, acc.data_ptr())
    ctx.finalize_device(acc.data_ptr(), 1.0 / 8, out.data_ptr())
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(64, 64, 3)
    assert np.allclose(got, whole, rtol=2e-5, atol=1e-6)  # float atomics: summation order only


def test_small_wavefront_capacity_gives_same_image(api, ctx, cornell):
    sc = cornell.with_size(48, 48)
    ctx.upload(sc)
    a = ctx.render_path(16, seed=6)
    ctx.configure(paths_in_flight=1000)  # forces many refill iterations
    b = ctx.render_path(16, seed=6)
    ctx.configure(0)
    assert np.allclose(a, b, rtol=2e-5, atol=1e-6)
    assert ctx.stats()["iterations"] >= 0


def test_no_lights_and_background(api, oracle, ctx, cornell):
    """Scene without emitters: sampleLight returns pdf 0 (IIntegrator.hpp:177-181); a miss at depth 0
    returns bkgcolor, a missed x_inter adds nothing (PathTracing.hpp:150,234)."""
    sc = cornell.with_size(32, 32)
    sc.materials = sc.materials.copy()
    sc.materials["emission"] = 0
    sc.bkgcolor = (0.2, 0.3, 0.4)
    ctx.upload(sc)
    g = ctx.render_path(4, seed=1)
    o = oracle.OracleScene(sc).render_path(4, seed=1)
    assert np.allclose(g, o, atol=1e-6)
    assert ((g == 0).all(-1) | np.isclose(g, np.array([0.2, 0.3, 0.4], np.float32)).all(-1)).all()


def test_render_errors(api, cornell):
    c = api.Context(0)
    with pytest.raises(api.TutuError):
        c.render_path(4)
    c.upload(cornell.with_size(8, 8))
    with pytest.raises(api.TutuError):
        c.render_path(0)
    # pipeline knob: 0..2 only; the setting survives a failed call
    assert api.lib().tutu_render_pipeline(c._h, 3) != 0
    assert api.lib().tutu_render_pipeline(c._h, -1) != 0
    assert np.isfinite(c.render_path(1)).all()
    c.close()


def _cornell_with_spheres(api, cornell):
    """The Cornell shell and light (12 triangles) with the two blocks replaced by four spheres: a scene of
    <= 32 primitives that has spheres, distinct and shared leaf boxes."""
    import copy
    sc = copy.copy(cornell)
    keep = cornell.prims[:12].copy()
    sph = np.zeros(4, api.PRIM_DTYPE)
    sph["tex_diffuse"] = sph["tex_normal"] = sph["tex_roughness"] = sph["tex_metallic"] = -1
    sph["type"] = api.PRIM_SPHERE
    sph["material"] = cornell.prims["material"][[12, 22, 12, 22]]
    for k, (c, r) in enumerate((((186, 82.5, 169), 82.5), ((368, 120, 351), 120), ((420, 60, 120), 60), ((120, 40, 420), 40))):
        sph["v"][k, 0:3] = c
        sph["v"][k, 3] = r
    sc.prims = np.concatenate([keep, sph])
    sc.bvh_nodes = None  # built on upload (midpoint rule)
    return sc


def test_small_scene_with_spheres_same_stream(api, oracle, ctx, cornell):
    """The small-scene kernels (one slab test per distinct leaf box, first pass of two candidates, parked
    rays finished 32 at a time — trace.cuh SmallPark) against the oracle on the same random numbers."""
    sc = _cornell_with_spheres(api, cornell).with_size(72, 72)
    ctx.upload(sc)
    g = ctx.render_path(12, seed=31)
    o, cnt = oracle.OracleScene(sc).render_path(12, seed=31, counters=True)
    d = np.abs(g - o)
    # <= 3 % of the pixels (12 paths each) hold a path that took another branch: the spheres fill a third
    # of the frame and Sphere::intersect / the normalisations differ from libm in the last place.  The
    # tree walk (traversal mode 4) gives the same 2.3 % and the same ray counts as the flat kernels.
    assert (d > 1e-3 * (1 + np.abs(o))).any(-1).mean() < 0.03
    assert np.median(d) < 1e-6
    assert abs(g.mean() / o.mean() - 1) < 5e-3
    st = ctx.stats()
    assert cnt[0] * 0.998 <= st["extend_rays"] <= cnt[0] * 1.002
    assert st["shadow_rays"] <= cnt[1]


def test_pipelines_give_the_same_paths(api, ctx, cornell, mixed):
    """tutu_render_pipeline: the register-resident kernel (resident.cuh) runs the same shade_vertex and flat
    tests as the wavefront on the same random numbers.  The two are compiled in different translation
    units, so FMA contraction may differ in the last place: images agree to float noise except for a
    handful of paths that flip a branch (ray counts within 1e-4)."""
    # (the 640^2 render is large enough for wf_shade to append through its reserves and leave dead queue entries)
    for sc, spp in ((cornell.with_size(80, 80), 12), (_cornell_with_spheres(api, cornell).with_size(64, 64), 8),
                    (cornell.with_size(640, 640), 10)):
        ctx.upload(sc)
        ctx.pipeline("wavefront")
        a = ctx.render_path(spp, seed=8)
        sa = ctx.stats()
        ctx.pipeline("resident")
        b = ctx.render_path(spp, seed=8)
        sb = ctx.stats()
        ctx.pipeline("auto")
        assert sb["kernel_launches"] == 1 and sa["kernel_launches"] > 1
        assert sa["paths"] == sb["paths"] and sa["nan_samples"] == sb["nan_samples"]
        for k in ("extend_rays", "shadow_rays"):
            assert abs(sa[k] - sb[k]) <= 1e-4 * sa[k], k
        assert (np.abs(a - b) > 1e-4 * (1 + np.abs(a))).any(-1).mean() < 2e-3
        assert abs(a.mean() / b.mean() - 1) < 1e-4
    # sample ranges compose in the resident pipeline too
    import torch
    sc = cornell.with_size(40, 40)
    ctx.upload(sc)
    ctx.pipeline("resident")
    whole = ctx.render_path(6, seed=4)
    acc = torch.zeros(40 * 40 * 3, dtype=torch.float32, device="cuda")
    out = torch.empty_like(acc)
    torch.cuda.synchronize()
    ctx.render_accumulate_device(0, 2, 4, acc.data_ptr())
    ctx.render_accumulate_device(2, 4, 4, acc.data_ptr())
    ctx.finalize_device(acc.data_ptr(), 1.0 / 6, out.data_ptr())
    torch.cuda.synchronize()
    assert np.allclose(out.cpu().numpy().reshape(40, 40, 3), whole, rtol=2e-5, atol=1e-6)
    # a scene that does not fit the constant bank cannot use it
    ctx.upload(mixed.with_size(16, 16))
    with pytest.raises(api.TutuError):
        ctx.render_path(2)
    ctx.pipeline("auto")
    assert np.isfinite(ctx.render_path(2)).all()


def test_scene_switch_between_trees_of_different_depth(api, ctx, golden):
    """The cached launch configuration of the wavefront depends on the traversal-stack size: a context that
    rendered one tree must render a deeper one (regression: CUDA 'invalid argument' on the second scene)."""
    means = []
    for name in ("glass_c4", "veach_80x60", "mixed", "glass_c4"):
        ctx.upload(api.Scene.load(golden / f"{name}.tscene").with_size(24, 24))
        img = ctx.render_path(2, seed=3)
        assert np.isfinite(img).all()
        means.append(float(img.mean()))
    assert means[0] == pytest.approx(means[3], rel=1e-5)


def test_small_scene_kernels_equal_the_tree_walk_at_full_size(api, ctx, cornell):
    """BASELINE size (1024x1024): the flat small-scene kernels (distinct-box slab tests, parked rays) and the
    general stack walk over the same tree must find the same hit for every one of the 10^7 queued rays, so
    the two renders trace the same paths: equal ray counts, images equal up to the order of the float
    atomics."""
    sc = cornell.with_size(1024, 1024)
    ctx.upload(sc)
    a = ctx.render_path(2, seed=13)
    sa = ctx.stats()
    ctx.set_traversal_mode(4)  # the general tree walk instead of the small-scene kernels
    b = ctx.render_path(2, seed=13)
    sb = ctx.stats()
    ctx.set_traversal_mode(0)
    for k in ("paths", "extend_rays", "shadow_rays", "nan_samples"):
        assert sa[k] == sb[k], k
    assert np.allclose(a, b, rtol=2e-5, atol=1e-6)

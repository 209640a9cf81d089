"""Parity tests for the BDPT integrator (tutu_render_bdpt, reference include/BDPT.hpp), through the
C ABI on a B200.

* same-stream: the CPU oracle (oracle/tutu_oracle_bdpt.hpp) draws the same Philox slots, so both
  build the same eye / light sub-paths and weigh the same (s,t) strategies;
* statistical: against high-spp means of the UNMODIFIED reference's BDPT (tests/golden/*_bdpt_*,
  tests/tools/make_golden.py --bdpt) on the Cornell box and on the Veach room of
  src/main_veach_bdpt.cpp (BASELINE.json configs[4]).
Tolerances are written next to each assertion."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rmse(a, b):
    return float(np.sqrt(((a - b) ** 2).mean()))


@pytest.mark.parametrize("name,w,h,spp", [("cornell_256", 48, 48, 8), ("veach_80x60", 40, 30, 8), ("mixed", 48, 48, 4)])
def test_bdpt_same_stream_as_oracle(api, oracle, ctx, golden, name, w, h, spp):
    sc = api.Scene.load(golden / f"{name}.tscene").with_size(w, h)
    ctx.upload(sc)
    g = ctx.render_bdpt(spp, seed=21)
    o, cnt = oracle.OracleScene(sc).render_bdpt(spp, seed=21, counters=True)
    assert np.isfinite(g).all() and np.isfinite(o).all()
    d = np.abs(g - o)
    # <= 2 % of the pixels may differ visibly: discrete flips (libm vs CUDA ulps at a branch, a
    # t = 1 splat landing one pixel over); BDPT's weights are ratios of products of pdfs
    assert (d > 1e-3 * (1 + np.abs(o))).any(-1).mean() < 0.02
    assert np.median(d) < 1e-5
    st = ctx.stats()
    assert st["paths"] == w * h * spp
    # closest-hit rays agree up to bookkeeping: the reference traces one more ray after the last
    # storable vertex (BDPT.hpp:286 runs before the loop condition :236) and stops sampling a pixel
    # whose primary ray missed (:733), the GPU does neither
    assert cnt[0] * 0.85 <= st["extend_rays"] <= cnt[0] * 1.15
    # any-hit rays: the reference tests visibility before it knows the contribution is zero
    assert st["shadow_rays"] <= cnt[1]


def test_bdpt_cornell_statistics_against_reference(api, ctx, golden, cornell):
    stats = json.loads((golden / "stats.json").read_text())["cornell_64_bdpt"]
    ref = np.fromfile(golden / "cornell_64_bdpt_ref_mean_4096.f32", np.float32).reshape(64, 64, 3)
    ctx.upload(cornell.with_size(64, 64))
    # noise level at 16 spp within +-10 % of the reference's own
    vals = [_rmse(ctx.render_bdpt(16, seed=s), ref) for s in (1, 2, 3, 4)]
    assert abs(np.mean(vals) / stats["rmse_16"] - 1) < 0.10
    # bias: 8192 spp vs the reference's 4096-spp mean
    big = ctx.render_bdpt(8192, seed=5)
    for c in range(3):
        assert abs(big[..., c].mean() / stats["channel_means"][c] - 1) < 0.005
    assert _rmse(big, ref) < 2 * stats["run_to_run_rmse_2048"]


def test_bdpt_veach_statistics_against_reference(api, ctx, golden):
    stats = json.loads((golden / "stats.json").read_text())["veach_80x60_bdpt"]
    ref = np.fromfile(golden / "veach_80x60_bdpt_ref_mean_2048.f32", np.float32).reshape(60, 80, 3)
    ctx.upload(api.Scene.load(golden / "veach_80x60.tscene"))
    vals = [_rmse(ctx.render_bdpt(16, seed=s), ref) for s in (1, 2, 3, 4)]
    assert abs(np.mean(vals) / stats["rmse_16"] - 1) < 0.15  # caustic fireflies make this noisier than Cornell
    big = ctx.render_bdpt(4096, seed=5)
    for c in range(3):
        assert abs(big[..., c].mean() / stats["channel_means"][c] - 1) < 0.01
    assert _rmse(big, ref) < 2 * stats["run_to_run_rmse_1024"]


def test_bdpt_glass_texture_scene_robust_statistics(api, oracle, ctx, golden):
    """MICROFACET_T / MICROFACET_R / all four texture channels through buildEyePath, buildLightPath
    (adjoint BSDF) and MISweight: the configs[3] stand-in scene under BDPT.  The reference's BDPT
    produces unbounded fireflies here (image mean 7.5 where path tracing gives 0.375), so the gates
    are robust statistics at EQUAL spp (tests/golden/stats.json: four reference runs at 256 spp):
    means of the image clipped at 2 and per-channel medians within 5 % (the reference's own runs
    spread by 3 %), plus the same-stream comparison with the oracle."""
    stats = json.loads((golden / "stats.json").read_text())["glass_c4_64_bdpt"]
    sc = api.Scene.load(golden / "glass_c4.tscene").with_size(64, 64)
    ctx.upload(sc)
    imgs = [ctx.render_bdpt(256, seed=s) for s in (1, 2, 3, 4)]
    cm = np.mean([np.clip(a, 0, 2).mean((0, 1)) for a in imgs], 0)
    md = np.mean([np.median(a, (0, 1)) for a in imgs], 0)
    assert np.isfinite(np.stack(imgs)).all()
    assert np.all(np.abs(cm / np.array(stats["spp256_clipped2_channel_means"]) - 1) < 0.05)
    assert np.all(np.abs(md / np.array(stats["spp256_channel_medians"]) - 1) < 0.05)
    small = sc.with_size(32, 32)
    ctx.upload(small)
    g = ctx.render_bdpt(8, seed=11)
    o = oracle.OracleScene(small).render_bdpt(8, seed=11)
    assert (np.abs(g - o) > 1e-3 * (1 + np.abs(o))).any(-1).mean() < 0.03
    assert np.median(np.abs(g - o)) < 1e-5


def test_bdpt_sample_ranges_compose_and_background(api, ctx, cornell):
    """[0,6) in one call == [0,2) + [2,6) accumulated; bkgcolor is added once by the finalize."""
    import torch
    sc = cornell.with_size(40, 40)
    sc.bkgcolor = (0.1, 0.2, 0.3)
    ctx.upload(sc)
    whole = ctx.render_bdpt(6, seed=4)
    acc = torch.zeros(40 * 40 * 3, dtype=torch.float32, device="cuda")
    out = torch.empty_like(acc)
    torch.cuda.synchronize()
    ctx.render_bdpt_accumulate_device(0, 2, 4, acc.data_ptr())
    ctx.render_bdpt_accumulate_device(2, 4, 4, acc.data_ptr())
    ctx.finalize_bdpt_device(acc.data_ptr(), 1.0 / 6, out.data_ptr())
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(40, 40, 3)
    assert np.allclose(got, whole, rtol=2e-5, atol=1e-6)
    # the open front of the box: primary rays that miss keep bkgcolor unless a t = 1 splat lands there
    assert (got >= np.array([0.1, 0.2, 0.3], np.float32) - 1e-6).all()


def test_bdpt_small_batches_give_same_image(api, ctx, cornell):
    sc = cornell.with_size(32, 32)
    ctx.upload(sc)
    a = ctx.render_bdpt(8, seed=6)
    ctx.configure(paths_in_flight=1000)  # 9 batches instead of 1
    b = ctx.render_bdpt(8, seed=6)
    ctx.configure(0)
    assert np.allclose(a, b, rtol=2e-5, atol=1e-6)


def test_bdpt_errors(api, cornell):
    c = api.Context(0)
    with pytest.raises(api.TutuError):
        c.render_bdpt(4)
    c.upload(cornell.with_size(8, 8))
    with pytest.raises(api.TutuError):
        c.render_bdpt(0)
    c.close()

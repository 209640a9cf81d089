"""The tree kernels keep their traversal stack in local memory for small trees and in shared memory for big ones
(tutu_traversal_stack).  The golden scenes are all small and the full-size scene is big, so the automatic
choice alone would leave half of the kernel variants untested: every case here runs with both flavours forced."""
import numpy as np
import pytest

from conftest import assert_hits_equal, load_rays
from test_gpu_trace import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("stack", ["shared", "local"])
@pytest.mark.parametrize("scene,rays,closest,anyf", [c for c in CASES if c[0] != "cornell_256"])
def test_golden_vectors_bit_exact_with_either_stack(api, ctx, golden, scene, rays, closest, anyf, stack):
    ctx.traversal_stack(stack)
    ctx.upload(api.Scene.load(golden / f"{scene}.tscene"))
    r = load_rays(golden / rays)
    for mode in (0, 3):
        ctx.set_traversal_mode(mode)
        assert_hits_equal(ctx.trace_closest(r), np.fromfile(golden / closest, api.HIT_DTYPE))
        assert np.array_equal(ctx.trace_any(r), np.fromfile(golden / anyf, np.uint8))


@pytest.mark.parametrize("stack", ["shared", "local"])
@pytest.mark.parametrize("name,size,spp", [("mixed", 96, 8), ("glass_c4", 128, 6)])
def test_wavefront_same_stream_as_oracle_with_either_stack(api, oracle, ctx, golden, name, size, spp, stack):
    """wf_extend / wf_shadow <0> (shared) and <2> (local) against the oracle on the same Philox stream (gates of
    test_gpu_render.py::test_same_stream_as_oracle); ray counts of the two flavours are equal."""
    sc = api.Scene.load(golden / f"{name}.tscene").with_size(size, size)
    ctx.traversal_stack(stack)
    ctx.upload(sc)
    g = ctx.render_path(spp, seed=33)
    st = ctx.stats()
    o = oracle.OracleScene(sc).render_path(spp, seed=33)
    d = np.abs(g - o)
    assert (d > 1e-3 * (1 + np.abs(o))).any(-1).mean() < 0.02  # discrete flips: libm vs CUDA ulps at branch points
    assert abs(g.mean() / o.mean() - 1) < 5e-3
    ctx.traversal_stack("local" if stack == "shared" else "shared")
    g2 = ctx.render_path(spp, seed=33)
    st2 = ctx.stats()
    assert (st["extend_rays"], st["shadow_rays"]) == (st2["extend_rays"], st2["shadow_rays"])
    assert np.allclose(g, g2, rtol=1e-4, atol=1e-6)  # the same paths; sums of atomics in another order


@pytest.mark.parametrize("stack", ["shared", "local"])
def test_bdpt_same_stream_as_oracle_with_either_stack(api, oracle, ctx, golden, stack):
    """q_extend / q_shadow_add <0> and <2> on the Veach room against the oracle's BDPT on the same stream."""
    sc = api.Scene.load(golden / "veach_80x60.tscene").with_size(40, 30)
    ctx.traversal_stack(stack)
    ctx.upload(sc)
    g = ctx.render_bdpt(8, seed=21)
    o = oracle.OracleScene(sc).render_bdpt(8, seed=21)
    d = np.abs(g - o)
    assert (d > 1e-3 * (1 + np.abs(o))).any(-1).mean() < 0.02  # the gates of test_gpu_bdpt.py::test_bdpt_same_stream_as_oracle
    assert np.median(d) < 1e-5


def test_full_size_any_hit_identical_with_either_stack(api, ctx):
    """configs[1]: k_trace_any<0> (shared; what the 999 698-triangle scene gets) and <4> (local) agree on 2^21 rays of each kind,
    and so do the closest hits (whose batch kernel keeps the shared stack)."""
    prims = api.synth_heightfield(707)
    sc = api.Scene(prims=prims, materials=api.default_material(), bvh_nodes=api.bvh_build(prims))
    ctx.upload(sc)
    for kind in (0, 1):
        rays = api.synth_rays(kind, 1 << 21, first=3 << 21)
        res = {}
        for stack in ("auto", "shared", "local"):
            ctx.traversal_stack(stack)
            res[stack] = (ctx.trace_any(rays), ctx.trace_closest(rays))
        assert np.array_equal(res["auto"][0], res["shared"][0]) and np.array_equal(res["auto"][0], res["local"][0])
        assert_hits_equal(res["auto"][1], res["local"][1])
        blocked = res["auto"][0].mean()
        assert 0.01 < blocked < 0.99  # both answers occur (97 % of the top-down rays are blocked)


@pytest.mark.parametrize("stack", ["shared", "local"])
@pytest.mark.parametrize("tracer", ["packets", "lanes"])
@pytest.mark.parametrize("name,w,h,spp", [("veach_80x60", 40, 30, 8), ("mixed", 48, 48, 4)])
def test_bdpt_queue_tracers_same_stream_as_oracle(api, oracle, ctx, golden, name, w, h, spp, tracer, stack):
    """q_extend / q_shadow_add as packet tracers and as persistent-lane tracers (tutu_bdpt_queue_tracer), with either stack,
    against the oracle's BDPT on the same stream; gates of test_gpu_bdpt.py::test_bdpt_same_stream_as_oracle."""
    sc = api.Scene.load(golden / f"{name}.tscene").with_size(w, h)
    ctx.traversal_stack(stack)
    ctx.bdpt_queue_tracer(tracer)
    ctx.upload(sc)
    g = ctx.render_bdpt(spp, seed=21)
    st = ctx.stats()
    o, cnt = oracle.OracleScene(sc).render_bdpt(spp, seed=21, counters=True)
    d = np.abs(g - o)
    assert (d > 1e-3 * (1 + np.abs(o))).any(-1).mean() < 0.02
    assert np.median(d) < 1e-5
    assert cnt[0] * 0.85 <= st["extend_rays"] <= cnt[0] * 1.15 and st["shadow_rays"] <= cnt[1]


def test_bdpt_queue_tracer_is_measured_per_scene(api, ctx, golden):
    """The automatic choice: a render of >= 8 batches runs one batch with each tracer and keeps the faster one until the
    next upload; the image is the same whichever wins (same sample set; sums of atomics in another order)."""
    sc = api.Scene.load(golden / "veach_80x60.tscene").with_size(400, 300)
    ctx.upload(sc)
    ctx.configure(1 << 20, False, 0)  # batches of 2^20 samples (default 2^22)
    assert ctx.bdpt_queue_tracer_measured()["tracer"] is None
    ctx.render_bdpt(2, seed=1)  # 240 000 samples = one batch: too short to measure
    assert ctx.bdpt_queue_tracer_measured()["tracer"] is None
    img = ctx.render_bdpt(80, seed=2)  # 9.6 M samples = 10 batches of 2^20
    m = ctx.bdpt_queue_tracer_measured()
    assert m["tracer"] in ("packets", "lanes") and m["ms_packets"] > 0 and m["ms_lanes"] > 0
    assert (m["tracer"] == "lanes") == (m["ms_lanes"] < m["ms_packets"])
    for forced in ("packets", "lanes"):
        ctx.bdpt_queue_tracer(forced)
        other = ctx.render_bdpt(80, seed=2)
        assert np.allclose(img, other, rtol=2e-3, atol=1e-5)
    ctx.bdpt_queue_tracer("auto")
    ctx.upload(sc)
    assert ctx.bdpt_queue_tracer_measured()["tracer"] is None

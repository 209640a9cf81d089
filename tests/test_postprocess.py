"""Output stage: the reference's Postprocessor (Postprocessor.hpp:29-197) — emissive extract,
separable Gaussian blur, add, exposure tone map — against images produced by the UNMODIFIED reference
(tests/golden/post_ref_*, written by tests/tools/make_golden.py --post through ref_harness postprocess)."""
import numpy as np
import pytest

CASES = [("48x40", (40, 48), ("extract", "blur", "bloom", "hdr", "full")), ("1030x8", (8, 1030), ("blur", "full"))]


def _load(golden, tag, shape, mode=None):
    name = f"post_in_{tag}.f32" if mode is None else f"post_ref_{tag}_{mode}.f32"
    return np.fromfile(golden / name, np.float32).reshape(*shape, 3)


def _bits_equal(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("tag,shape,modes", CASES)
def test_oracle_postprocess_equals_reference(oracle, golden, tag, shape, modes):
    src = _load(golden, tag, shape)
    for mode in modes:
        assert _bits_equal(oracle.postprocess(src, mode), _load(golden, tag, shape, mode)), mode


def test_fixture_exercises_the_texture_fetch_quirks(golden):
    """Texture::getRGBat maps u == 0 to 1 (Texture.hpp:24-26): row 0 of every stage reads the LAST texel and
    column 0 reads the first texel of the next row.  The fixture must show it, or the gate would not."""
    src = _load(golden, "48x40", (40, 48))
    ext = _load(golden, "48x40", (40, 48), "extract")
    assert np.array_equal(ext[0, 5], [0, 2, 0])      # row 0 <- last texel (0, 9, 0), rescaled to strength 2
    assert np.array_equal(ext[4, 0], [2, 0, 0])      # column 0 of row 4 <- texel (5, 0) = (4, 0, 0)
    assert (ext[10:14, 20:26] == np.float32([2, 1.6, 1.2])).all()
    assert src[0, 0, 0] == 5 and ext[0, 0, 1] == 2   # its own value is never read


@pytest.mark.gpu
@pytest.mark.parametrize("tag,shape,modes", CASES)
def test_gpu_postprocess_equals_reference(api, ctx, golden, tag, shape, modes):
    src = _load(golden, tag, shape)
    for mode in modes:
        got, want = ctx.postprocess(src, mode), _load(golden, tag, shape, mode)
        if mode in ("hdr", "full"):
            # 1 - expf(-c * 1.5): glibc's expf vs exp() in double rounded once; both within 1 ulp of exp(),
            # so the outputs differ by at most one ulp of 1.0 and agree exactly almost everywhere
            assert np.abs(got - want).max() <= 1.2e-7, mode
            assert (got.view(np.uint32) == want.view(np.uint32)).mean() > 0.995, mode
        else:
            assert _bits_equal(got, want), mode


@pytest.mark.gpu
def test_gpu_postprocess_large_and_custom_parameters(api, oracle, ctx):
    rng = np.random.default_rng(3)
    img = rng.uniform(0, 1.5, (331, 257, 3)).astype(np.float32)
    img[rng.random((331, 257)) < 0.02] *= 8.0
    for mode in ("extract", "blur", "bloom"):
        assert _bits_equal(ctx.postprocess(img, mode), oracle.postprocess(img, mode)), mode
    assert np.abs(ctx.postprocess(img, "full") - oracle.postprocess(img, "full")).max() <= 1.2e-7
    # other constants than the reference's #defines: odd kernel (taps -7..7), two extra blur loops
    p = api.post_params(kernel_size=15, stddev=4.0, gaussian_loops=2, strength=1.25, emissive_norm=2.0, exposure=0.7)
    assert _bits_equal(ctx.postprocess(img, "bloom", p), oracle.postprocess(img, "bloom", p))
    assert np.abs(ctx.postprocess(img, "full", p) - oracle.postprocess(img, "full", p)).max() <= 1.2e-7
    one = img[:1, :1].copy()  # 1 x 1 image: every tap clamps onto the only texel
    assert _bits_equal(ctx.postprocess(one, "bloom"), oracle.postprocess(one, "bloom"))


@pytest.mark.gpu
def test_gpu_postprocess_device_entry_and_errors(api, oracle, ctx):
    import torch
    rng = np.random.default_rng(4)
    img = rng.uniform(0, 4, (96, 128, 3)).astype(np.float32)
    d_in = torch.from_numpy(img).cuda()
    d_out = torch.empty_like(d_in)
    torch.cuda.synchronize()
    ctx.postprocess_device(d_in.data_ptr(), 128, 96, d_out.data_ptr(), "bloom",
                           stream=api.stream_handle(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert _bits_equal(d_out.cpu().numpy(), oracle.postprocess(img, "bloom"))
    with pytest.raises(api.TutuError):  # in place is not possible: every output pixel gathers other texels
        ctx.postprocess_device(d_in.data_ptr(), 128, 96, d_in.data_ptr(), "hdr")
    with pytest.raises(api.TutuError):
        ctx.postprocess(img, "blur", api.post_params(kernel_size=65))
    with pytest.raises(api.TutuError):
        ctx.postprocess(img, "blur", api.post_params(stddev=0.0))

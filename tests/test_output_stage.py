"""Output stage: PPMGenerator::writePixel's gamma-0.78 quantisation and the
ASCII P3 file, against a PPM written by the UNMODIFIED reference (tests/golden/ppm_ref_64x48.ppm,
tests/tools/make_golden.py --ppm)."""
import numpy as np
import pytest


def _ref_pixels(golden):
    tok = (golden / "ppm_ref_64x48.ppm").read_bytes().split()
    assert tok[:4] == [b"P3", b"64", b"48", b"255"]
    return np.array(tok[4:], dtype=np.int64).reshape(48, 64, 3).astype(np.uint8)


def _img(golden):
    return np.fromfile(golden / "ppm_in_64x48.f32", np.float32).reshape(48, 64, 3)


def test_oracle_write_pixel_equals_reference_ppm(oracle, golden):
    assert np.array_equal(oracle.write_pixel(_img(golden)), _ref_pixels(golden))


def test_write_ppm_reproduces_reference_file_byte_for_byte(api, golden, tmp_path):
    out = tmp_path / "o.ppm"
    api.write_ppm(out, _ref_pixels(golden), binary=False)
    assert out.read_bytes() == (golden / "ppm_ref_64x48.ppm").read_bytes()
    api.write_ppm(out, _ref_pixels(golden), binary=True)
    raw = out.read_bytes()
    assert raw.startswith(b"P6\n64 48\n255\n") and raw[-64 * 48 * 3:] == _ref_pixels(golden).tobytes()


@pytest.mark.gpu
def test_gpu_quantize_equals_reference_ppm(api, oracle, ctx, golden):
    got = ctx.quantize(_img(golden))
    assert np.array_equal(got, _ref_pixels(golden))  # bit-exact incl. NaN / inf / negative / > 1 pixels
    # a large random image against the oracle; odd size exercises the tail
    rng = np.random.default_rng(1)
    big = rng.uniform(-0.2, 1.3, (331, 257, 3)).astype(np.float32)
    assert np.array_equal(ctx.quantize(big), oracle.write_pixel(big))
    assert np.array_equal(ctx.quantize(big, gamma=0.0), oracle.write_pixel(big, 0.0))


@pytest.mark.gpu
def test_render_to_ppm_end_to_end(api, oracle, ctx, cornell, tmp_path):
    """render -> quantise on the device -> P3 file == the oracle's quantisation of the same image."""
    import torch
    sc = cornell.with_size(64, 64)
    ctx.upload(sc)
    acc = torch.zeros(64 * 64 * 3, dtype=torch.float32, device="cuda")
    rgb = torch.empty_like(acc)
    q = torch.empty(64 * 64 * 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.render_accumulate_device(0, 8, 3, acc.data_ptr())
    ctx.finalize_device(acc.data_ptr(), 1.0 / 8, rgb.data_ptr())
    ctx.quantize_device(rgb.data_ptr(), 64 * 64, q.data_ptr())
    torch.cuda.synchronize()
    img8 = q.cpu().numpy().reshape(64, 64, 3)
    assert np.array_equal(img8, oracle.write_pixel(rgb.cpu().numpy().reshape(64, 64, 3)))
    api.write_ppm(tmp_path / "cornell.ppm", img8)
    assert (tmp_path / "cornell.ppm").read_bytes().startswith(b"P3\n64\n64\n255\n")

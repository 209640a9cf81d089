"""f-2, texture files on the host side of the ABI: tutu_texture_load reads ASCII P3 exactly as the
reference's PPMGenerator::loadTexture does (PPMGenerator.hpp:1027-1084, checked against the reference itself
through ref_harness load-texture) and additionally binary P6 and PNG, which the reference cannot read (:1050)."""
import numpy as np
import pytest


def _img(h=13, w=17, seed=0, maxval=255):
    rng = np.random.default_rng(seed)
    return rng.integers(0, maxval + 1, (h, w, 3)).astype(np.uint16 if maxval > 255 else np.uint8)


def test_p3_equals_the_reference_loader(api, tmp_path):
    from oracle import oracle_py as O
    assert O.ref_available(), "oracle/_ref/ref_harness is missing"
    img = _img()
    img[0, :8] = [[0, 0, 0], [255, 255, 255], [1, 2, 3], [254, 127, 128], [85, 170, 51], [17, 34, 68], [3, 5, 7], [250, 251, 252]]
    p = tmp_path / "t.ppm"
    api.write_ppm(p, img)  # ASCII P3, the reference's own format
    mine = api.load_texture(p)
    info = O._run(["load-texture", str(p), str(tmp_path / "ref.f32")])
    ref = np.fromfile(tmp_path / "ref.f32", np.float32).reshape(info["height"], info["width"], 3)
    assert mine.shape == (13, 17, 3) and np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
    assert np.array_equal(mine, img.astype(np.float32) / np.float32(255))
    # bump maps: c * 2 - 1 as PPMGenerator.hpp:714-720
    nm = api.load_texture(p, normal_map=True)
    assert np.array_equal(nm, mine * np.float32(2) - np.float32(1))


def test_p6_and_png_give_the_same_texels_as_p3(api, tmp_path):
    img = _img(seed=2)
    api.write_ppm(tmp_path / "a.ppm", img)
    api.write_ppm(tmp_path / "b.ppm", img, binary=True)
    api.write_png(tmp_path / "c.png", img)
    a, b, c = (api.load_texture(tmp_path / n) for n in ("a.ppm", "b.ppm", "c.png"))
    assert np.array_equal(a, b) and np.array_equal(a, c)
    # header comments and 16-bit P6
    raw = (tmp_path / "b.ppm").read_bytes()
    (tmp_path / "d.ppm").write_bytes(raw.replace(b"P6\n", b"P6\n# made by a paint program\n", 1))
    assert np.array_equal(api.load_texture(tmp_path / "d.ppm"), a)
    wide = _img(5, 4, 3, maxval=65535)
    (tmp_path / "e.ppm").write_bytes(b"P6\n4 5\n65535\n" + wide.astype(">u2").tobytes())
    assert np.array_equal(api.load_texture(tmp_path / "e.ppm"), wide.astype(np.float32) / np.float32(65535))


def test_png_written_by_another_encoder(api, tmp_path):
    """PNG files from PIL: RGB, RGBA, grey, grey+alpha, palette, 16-bit grey, 1-bit; PIL picks its own scanline
    filters (Sub / Up / Average / Paeth all occur on noise-free gradients)."""
    Image = pytest.importorskip("PIL.Image")
    y, x = np.mgrid[0:37, 0:53]
    rgb = np.stack([(x * 4) % 256, (y * 6) % 256, (x * y) % 256], -1).astype(np.uint8)
    want = rgb.astype(np.float32) / np.float32(255)
    Image.fromarray(rgb).save(tmp_path / "rgb.png", optimize=True)
    assert np.array_equal(api.load_texture(tmp_path / "rgb.png"), want)
    rgba = np.concatenate([rgb, np.full((37, 53, 1), 77, np.uint8)], -1)
    Image.fromarray(rgba).save(tmp_path / "rgba.png")
    assert np.array_equal(api.load_texture(tmp_path / "rgba.png"), want)
    grey = rgb[..., 0]
    Image.fromarray(grey).save(tmp_path / "l.png")
    assert np.array_equal(api.load_texture(tmp_path / "l.png"), np.repeat(want[..., :1], 3, -1))
    Image.fromarray(np.stack([grey, 255 - grey], -1), "LA").save(tmp_path / "la.png")
    assert np.array_equal(api.load_texture(tmp_path / "la.png"), np.repeat(want[..., :1], 3, -1))
    pal = Image.fromarray(rgb).quantize(16)
    pal.save(tmp_path / "p.png")
    assert np.array_equal(api.load_texture(tmp_path / "p.png"), np.asarray(pal.convert("RGB"), np.float32) / np.float32(255))
    g16 = ((x * 1000 + y) % 65536).astype(np.uint16)
    Image.fromarray(g16).save(tmp_path / "l16.png")
    assert np.array_equal(api.load_texture(tmp_path / "l16.png")[..., 0], g16.astype(np.float32) / np.float32(65535))
    bits = ((x + y) % 2).astype(bool)
    Image.fromarray(bits).save(tmp_path / "l1.png")
    assert np.array_equal(api.load_texture(tmp_path / "l1.png")[..., 1], bits.astype(np.float32))
    # and our writer is readable by that decoder
    api.write_png(tmp_path / "ours.png", rgb)
    assert np.array_equal(np.asarray(Image.open(tmp_path / "ours.png")), rgb)


def test_texture_file_errors(api, tmp_path):
    with pytest.raises(api.TutuError):
        api.load_texture(tmp_path / "missing.ppm")
    (tmp_path / "x.ppm").write_bytes(b"P3\n2 2\n255\n1 2 3 4 5 6\n")  # truncated
    with pytest.raises(api.TutuError):
        api.load_texture(tmp_path / "x.ppm")
    (tmp_path / "y.ppm").write_bytes(b"P3\n2 1\n255\n1 2 3 -4 5 6\n")  # checkPosInt rejects signs (global.hpp:71-85)
    with pytest.raises(api.TutuError):
        api.load_texture(tmp_path / "y.ppm")
    (tmp_path / "z.bin").write_bytes(b"GIF89a....")
    with pytest.raises(api.TutuError):
        api.load_texture(tmp_path / "z.bin")
    img = _img(4, 4)
    api.write_png(tmp_path / "ok.png", img)
    raw = bytearray((tmp_path / "ok.png").read_bytes())
    raw[45] ^= 0xFF  # flip a byte inside IDAT: the chunk CRC catches it
    (tmp_path / "bad.png").write_bytes(raw)
    with pytest.raises(api.TutuError):
        api.load_texture(tmp_path / "bad.png")
    (tmp_path / "cut.png").write_bytes(bytes(raw[:40]))
    with pytest.raises(api.TutuError):
        api.load_texture(tmp_path / "cut.png")

"""ctypes binding of the CPU oracle (oracle/libtutu_oracle.so) and of the compiled reference
harness (oracle/_ref/ref_harness).

TEST INFRASTRUCTURE ONLY: imported by tests/, tools/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under tuturenderer_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libtutu_oracle.so"
REF_DIR = HERE / "_ref"
REF_HARNESS = REF_DIR / "ref_harness"

_lib = None


def build(ref: bool = True) -> None:
    """Compiles the oracle port and (when /root/reference is present) the reference harness."""
    subprocess.run(["make", "-s", "-C", str(HERE), "port"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build(ref=False)
        l = C.CDLL(str(LIB_PATH))
        P = C.c_void_p
        l.oracle_scene_create.restype = P
        l.oracle_scene_create.argtypes = [P]
        l.oracle_scene_destroy.argtypes = [P]
        l.oracle_bvh_node_count.restype = C.c_uint32
        l.oracle_bvh_node_count.argtypes = [P]
        l.oracle_bvh_export.restype = C.c_uint32
        l.oracle_bvh_export.argtypes = [P, P]
        l.oracle_trace_closest.argtypes = [P, P, C.c_uint64, P, C.c_int]
        l.oracle_trace_any.argtypes = [P, P, C.c_uint64, P, C.c_int]
        l.oracle_render_path.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, P, P, C.c_int]
        l.oracle_render_bdpt.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, P, P, C.c_int]
        l.oracle_primary_rays.argtypes = [P, P]
        l.oracle_write_pixel.argtypes = [P, C.c_uint64, C.c_float, P]
        l.oracle_postprocess.restype = C.c_int
        l.oracle_postprocess.argtypes = [P, C.c_int, C.c_int, C.c_int, P, P]
        _lib = l
    return _lib


class OracleScene:
    """The reference's algorithm restated on the CPU, over a tuturenderer_b200.api.Scene."""

    def __init__(self, scene):
        from tuturenderer_b200 import api  # PODs / dtypes only
        self._api = api
        self.scene = scene
        d, keep = scene.to_c()
        self._h = lib().oracle_scene_create(C.byref(d))
        del keep
        if not self._h:
            raise RuntimeError("oracle_scene_create failed")

    def close(self):
        if getattr(self, "_h", None):
            lib().oracle_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def bvh_export(self) -> np.ndarray:
        n = lib().oracle_bvh_node_count(self._h)
        out = np.zeros(max(n, 1), self._api.BVHNODE_DTYPE)
        cnt = lib().oracle_bvh_export(self._h, out.ctypes.data)
        return out[:cnt]

    def trace_closest(self, rays: np.ndarray, threads: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        out = np.empty(len(rays), self._api.HIT_DTYPE)
        lib().oracle_trace_closest(self._h, rays.ctypes.data, len(rays), out.ctypes.data, threads)
        return out

    def trace_any(self, rays: np.ndarray, threads: int = 0) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        out = np.empty(len(rays), np.uint8)
        lib().oracle_trace_any(self._h, rays.ctypes.data, len(rays), out.ctypes.data, threads)
        return out

    def render_path(self, spp: int, seed: int = 1, sample_begin: int = 0, total_spp: int | None = None,
                    threads: int = 0, counters: bool = False):
        out = np.empty((self.scene.height, self.scene.width, 3), np.float32)
        cnt = np.zeros(3, np.uint64)
        lib().oracle_render_path(self._h, sample_begin, spp, total_spp or spp, seed, out.ctypes.data,
                                 cnt.ctypes.data, threads)
        return (out, cnt) if counters else out

    def render_bdpt(self, spp: int, seed: int = 1, sample_begin: int = 0, total_spp: int | None = None,
                    threads: int = 0, counters: bool = False):
        """BDPT::integrate restated (bkgcolor + added contributions, incl. t = 1 splats)."""
        out = np.empty((self.scene.height, self.scene.width, 3), np.float32)
        cnt = np.zeros(3, np.uint64)
        lib().oracle_render_bdpt(self._h, sample_begin, spp, total_spp or spp, seed, out.ctypes.data,
                                 cnt.ctypes.data, threads)
        return (out, cnt) if counters else out

    def primary_rays(self) -> np.ndarray:
        out = np.empty((self.scene.height * self.scene.width, 8), np.float32)
        lib().oracle_primary_rays(self._h, out.ctypes.data)
        return out


def write_pixel(rgb: np.ndarray, gamma: float = 0.78) -> np.ndarray:
    """PPMGenerator::writePixel's quantisation restated (oracle_write_pixel)."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    out = np.empty(rgb.shape, np.uint8)
    lib().oracle_write_pixel(rgb.ctypes.data, rgb.size, gamma, out.ctypes.data)
    return out


POST_MODES = {"extract": 1, "blur": 2, "bloom": 3, "hdr": 4, "full": 5}  # TUTU_POST_* of include/tutu_b200.h


def postprocess(rgb: np.ndarray, mode: str, params=None) -> np.ndarray:
    """The reference's Postprocessor restated (oracle_postprocess); `params` = a tuturenderer_b200.api.TutuPostParams
    or None for the reference's #define constants."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    out = np.empty_like(rgb)
    rc = lib().oracle_postprocess(rgb.ctypes.data, rgb.shape[1], rgb.shape[0], POST_MODES[mode],
                                  C.byref(params) if params is not None else None, out.ctypes.data)
    if rc != 0:
        raise ValueError(f"oracle_postprocess: bad mode {mode}")
    return out


# ------------------------------------------------------------------------------------------------
# the compiled reference (oracle/_ref/ref_harness)
# ------------------------------------------------------------------------------------------------
def ref_available() -> bool:
    return REF_HARNESS.exists() and os.access(REF_HARNESS, os.X_OK)


def _run(args: list[str], timeout: float | None = None) -> dict:
    res = subprocess.run([str(REF_HARNESS), *args], capture_output=True, text=True, timeout=timeout)
    if res.returncode != 0:
        raise RuntimeError(f"ref_harness {' '.join(args)} failed ({res.returncode}): {res.stderr[-2000:]}")
    last = [l for l in res.stdout.strip().splitlines() if l.startswith("{")]
    return json.loads(last[-1]) if last else {}


def ref_dump_cornell(width: int, height: int, out_path) -> dict:
    return _run(["dump-cornell", str(REF_DIR / "model"), str(width), str(height), str(out_path)])


def ref_dump_veach(width: int, height: int, out_path) -> dict:
    """src/main_veach_bdpt.cpp's scene (2308 triangles) with the reference-built BVH."""
    return _run(["dump-veach", str(REF_DIR / "model"), str(width), str(height), str(out_path)])


def ref_ppm(rgb: np.ndarray, out_path) -> None:
    """The reference's own PPMGenerator::generate (header + writePixel) on a float image (H, W, 3)."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    with tempfile.TemporaryDirectory() as td:
        ip = Path(td) / "in.f32"
        rgb.tofile(ip)
        _run(["ppm", str(rgb.shape[1]), str(rgb.shape[0]), str(ip), str(out_path)])


def ref_export_bvh(scene, out_path=None):
    """Runs Scene::initializeBVH of the reference on `scene`; returns the scene with its tree."""
    from tuturenderer_b200.api import Scene
    with tempfile.TemporaryDirectory() as td:
        src = Path(td) / "in.tscene"
        dst = Path(out_path) if out_path else Path(td) / "out.tscene"
        scene.save(src)
        _run(["export-bvh", str(src), str(dst)])
        return Scene.load(dst)


HIT_DTYPE = np.dtype([("prim", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])  # TutuHit, include/tutu_b200.h


def ref_trace(scene, rays: np.ndarray, kind: str = "closest", threads: int = 0, scene_path=None):
    """getIntersection / hasIntersection of the reference; returns (result, info).  `scene_path`: an
    already saved .tscene of `scene` (big scenes traced several times are written once)."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    with tempfile.TemporaryDirectory() as td:
        sp, rp, op = Path(td) / "s.tscene", Path(td) / "r.f32", Path(td) / "o.bin"
        if scene_path is not None:
            sp = Path(scene_path)
        else:
            scene.save(sp)
        rays.tofile(rp)
        info = _run(["trace", str(sp), str(rp), kind, str(op), str(threads)])
        out = np.fromfile(op, dtype=np.uint8 if kind == "any" else HIT_DTYPE)
    return out, info


def ref_postprocess(rgb: np.ndarray, mode: str):
    """The reference's Postprocessor (Postprocessor.hpp:29-197, its own #define constants) on a linear float
    image (H, W, 3).  mode: "extract", "blur", "bloom", "hdr" or "full" (= performPostProcess under HDR_BLOOM).
    Returns (image, info)."""
    rgb = np.ascontiguousarray(rgb, np.float32)
    h, w = rgb.shape[:2]
    with tempfile.TemporaryDirectory() as td:
        ip, op = Path(td) / "in.f32", Path(td) / "out.f32"
        rgb.tofile(ip)
        info = _run(["postprocess", mode, str(w), str(h), str(ip), str(op)])
        return np.fromfile(op, np.float32).reshape(h, w, 3), info


# TUTUSCN1 header (tuturenderer_b200/csrc/host_scene.cpp: FileHeader): magic[8], version, n_prims, n_materials,
# n_bvh_nodes, n_tex[4], TutuCamera{eye[3], viewdir[3], updir[3], hfov, width, height, parallel}, bkg[3], eta
_HDR_WIDTH_OFFSET = 8 + 4 + 12 + 16 + 36 + 4


def ref_render_file(scene_file, spp: int, mode: str = "rows", width: int | None = None, height: int | None = None,
                    timeout: float | None = None):
    """PathTracing::integrate of the reference on a .tscene FILE, optionally at another frame size
    (the header's width/height are patched in a scratch copy).  Uses nothing of the product package:
    this is what bench.py --impl reference runs."""
    raw = bytearray(Path(scene_file).read_bytes())
    if raw[:8] != b"TUTUSCN1":
        raise RuntimeError(f"{scene_file} is not a TUTUSCN1 file")
    wh = np.frombuffer(raw, np.int32, 2, _HDR_WIDTH_OFFSET)
    if width is not None:
        wh[0] = width
    if height is not None:
        wh[1] = height
    w, h = int(wh[0]), int(wh[1])
    with tempfile.TemporaryDirectory() as td:
        sp, op = Path(td) / "s.tscene", Path(td) / "o.f32"
        sp.write_bytes(raw)
        info = _run(["render", str(sp), str(spp), str(op), mode], timeout=timeout)
        img = np.fromfile(op, dtype=np.float32).reshape(h, w, 3)
    return img, info


def ref_render(scene, spp: int, mode: str = "rows", timeout: float | None = None):
    """PathTracing::integrate of the reference; returns (linear float image, info)."""
    with tempfile.TemporaryDirectory() as td:
        sp, op = Path(td) / "s.tscene", Path(td) / "o.f32"
        scene.save(sp)
        info = _run(["render", str(sp), str(spp), str(op), mode], timeout=timeout)
        img = np.fromfile(op, dtype=np.float32).reshape(scene.height, scene.width, 3)
    return img, info

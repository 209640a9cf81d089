"""ctypes binding of libtutu_b200.so (include/tutu_b200.h) for the tests, bench.py and tools.

This is plumbing, not the product: every call goes straight to the C ABI.  If the shared
library is missing or cannot be loaded the import of :func:`lib` raises — there is no Python or
CPU stand-in for any entry point.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libtutu_b200.so"

TUTU_OK = 0
TUTU_E_INVALID, TUTU_E_CUDA, TUTU_E_STATE, TUTU_E_IO, TUTU_E_NOMEM = -1, -2, -3, -4, -5
PRIM_TRIANGLE, PRIM_SPHERE = 0, 1
MAT_LAMBERTIAN, MAT_PERFECT_REFLECTIVE, MAT_PERFECT_REFRACTIVE, MAT_MICROFACET_R, MAT_MICROFACET_T, MAT_UNLIT = range(6)
TEX_DIFFUSE, TEX_NORMAL, TEX_ROUGHNESS, TEX_METALLIC = range(4)
RAY_FLOATS = 8

# numpy mirrors of the PODs (little-endian, packed exactly like the C structs)
MATERIAL_DTYPE = np.dtype([("diffuse", "<f4", (3,)), ("specular", "<f4", (3,)), ("emission", "<f4", (3,)),
                           ("type", "<i4"), ("alpha", "<f4"), ("eta", "<f4"), ("roughness", "<f4"),
                           ("metallic", "<f4")])
PRIM_DTYPE = np.dtype([("type", "<i4"), ("v", "<f4", (9,)), ("n", "<f4", (9,)), ("uv", "<f4", (6,)),
                       ("material", "<i4"), ("tex_active", "<i4"), ("tex_diffuse", "<i4"),
                       ("tex_normal", "<i4"), ("tex_roughness", "<i4"), ("tex_metallic", "<i4")])
BVHNODE_DTYPE = np.dtype([("left", "<i4"), ("right", "<i4"), ("prim", "<i4")])
HIT_DTYPE = np.dtype([("prim", "<i4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])
assert MATERIAL_DTYPE.itemsize == 56 and PRIM_DTYPE.itemsize == 124
assert BVHNODE_DTYPE.itemsize == 12 and HIT_DTYPE.itemsize == 16


class TutuTexture(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.POINTER(C.c_float))]


class TutuCamera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("viewdir", C.c_float * 3), ("updir", C.c_float * 3),
                ("hfov_deg", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("parallel_projection", C.c_int32)]


class TutuSceneDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("n_prims", C.c_uint32), ("prims", C.c_void_p),
                ("n_materials", C.c_uint32), ("n_bvh_nodes", C.c_uint32), ("materials", C.c_void_p),
                ("bvh_nodes", C.c_void_p), ("tex", C.POINTER(TutuTexture) * 4), ("n_tex", C.c_uint32 * 4),
                ("camera", TutuCamera), ("bkgcolor", C.c_float * 3), ("eta", C.c_float)]


class TutuSceneInfo(C.Structure):
    _fields_ = [("n_prims", C.c_uint32), ("n_nodes", C.c_uint32), ("n_inner", C.c_uint32),
                ("depth", C.c_uint32), ("n_lights", C.c_uint32), ("n_materials", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32), ("device_bytes", C.c_uint64),
                ("trav_nodes", C.c_uint32), ("trav_depth", C.c_uint32), ("trav_width", C.c_uint32),
                ("trav_node_bytes", C.c_uint32), ("trav_leaf_bytes", C.c_uint32), ("reserved", C.c_uint32)]


class TutuRenderStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("shade_calls", C.c_uint64), ("nan_samples", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("iterations", C.c_uint64), ("gpu_ms", C.c_float), ("extend_ms", C.c_float),
                ("shade_ms", C.c_float), ("shadow_ms", C.c_float), ("other_ms", C.c_float)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class TutuUploadStats(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("flatten_ms", C.c_float), ("tree_build_ms", C.c_float), ("h2d_ms", C.c_float),
                ("builder", C.c_int32), ("tree_depth", C.c_uint32)]


BUILDERS = {"auto": 0, "host_sah": 1, "device_lbvh": 2, "device_ploc": 3, "device_sah": 4}


class TutuTreeCheck(C.Structure):
    _fields_ = [("n_leaves", C.c_uint32), ("binary_nodes", C.c_uint32), ("binary_depth", C.c_uint32),
                ("wide_nodes", C.c_uint32), ("wide_depth", C.c_uint32), ("wide_children", C.c_uint64),
                ("violations", C.c_uint64)]


class TutuPostParams(C.Structure):
    _fields_ = [("emissive_norm", C.c_float), ("strength", C.c_float), ("gaussian_loops", C.c_int32),
                ("kernel_size", C.c_int32), ("stddev", C.c_float), ("exposure", C.c_float)]


POST_MODES = {"extract": 1, "blur": 2, "bloom": 3, "hdr": 4, "hdr_bloom": 5, "full": 5}

# every symbol include/tutu_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
ABI = {
    "tutu_abi_version": (C.c_int, []),
    "tutu_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "tutu_ctx_destroy": (None, [_P]),
    "tutu_last_error": (C.c_char_p, [_P]),
    "tutu_scene_upload": (C.c_int, [_P, C.POINTER(TutuSceneDesc)]),
    "tutu_scene_info": (C.c_int, [_P, C.POINTER(TutuSceneInfo)]),
    "tutu_scene_set_camera": (C.c_int, [_P, C.POINTER(TutuCamera)]),
    "tutu_scene_builder": (C.c_int, [_P, C.c_int]),
    "tutu_upload_stats": (C.c_int, [_P, C.POINTER(TutuUploadStats)]),
    "tutu_trace_closest": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "tutu_trace_any": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "tutu_trace_closest_device": (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    "tutu_trace_any_device": (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    "tutu_set_traversal_mode": (C.c_int, [_P, C.c_int]),
    "tutu_traversal_stack": (C.c_int, [_P, C.c_int]),
    "tutu_bdpt_queue_tracer": (C.c_int, [_P, C.c_int]),
    "tutu_bdpt_queue_tracer_measured": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "tutu_trace_count_visits": (C.c_int, [_P, _P, C.c_uint64, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tutu_render_path": (C.c_int, [_P, C.c_uint32, C.c_uint64, _P]),
    "tutu_render_path_accumulate_device": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint64, _P, _P]),
    "tutu_finalize_device": (C.c_int, [_P, _P, C.c_float, _P, _P]),
    "tutu_render_bdpt": (C.c_int, [_P, C.c_uint32, C.c_uint64, _P]),
    "tutu_render_bdpt_accumulate_device": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint64, _P, _P]),
    "tutu_finalize_bdpt_device": (C.c_int, [_P, _P, C.c_float, _P, _P]),
    "tutu_quantize": (C.c_int, [_P, _P, C.c_uint64, C.c_float, _P]),
    "tutu_quantize_device": (C.c_int, [_P, _P, C.c_uint64, C.c_float, _P, _P]),
    "tutu_post_params_default": (None, [C.POINTER(TutuPostParams)]),
    "tutu_postprocess": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(TutuPostParams), _P]),
    "tutu_postprocess_device": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(TutuPostParams), _P, _P]),
    "tutu_write_ppm": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, _P, C.c_int]),
    "tutu_write_png": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, _P]),
    "tutu_texture_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "tutu_texture_free": (None, [C.POINTER(C.c_float)]),
    "tutu_render_stats": (C.c_int, [_P, C.POINTER(TutuRenderStats)]),
    "tutu_render_configure": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_int]),
    "tutu_render_pipeline": (C.c_int, [_P, C.c_int]),
    "tutu_bvh_build": (C.c_int, [_P, C.c_uint32, _P, C.POINTER(C.c_uint32)]),
    "tutu_traversal_tree_check": (C.c_int, [C.POINTER(TutuSceneDesc), C.POINTER(TutuTreeCheck)]),
    "tutu_scene_file_load": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "tutu_scene_file_desc": (C.POINTER(TutuSceneDesc), [_P]),
    "tutu_scene_file_free": (None, [_P]),
    "tutu_scene_file_save": (C.c_int, [C.POINTER(TutuSceneDesc), C.c_char_p]),
    "tutu_synth_heightfield": (C.c_int, [C.c_uint32, C.c_uint64, _P]),
    "tutu_synth_rays": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, _P]),
    "tutu_debug_guard_check": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tutu_debug_guard_poke": (C.c_int, [C.c_uint32]),
}

_lib = None

# torch's default stream has the handle 0, which the C ABI reads as "the ctx's own stream";
# CUDA's explicit name for the legacy default stream is cudaStreamLegacy = 0x1.
CUDA_STREAM_LEGACY = 1


CTX_STREAM = "ctx"  # opt-in: the context's own non-blocking stream (NOT ordered with torch's streams)


def stream_handle(stream) -> int | None:
    """Maps a stream argument of the Context methods to the handle the C ABI expects: a torch `cuda_stream`
    integer (0 = the legacy default stream, where torch's default-stream work is ordered) or CTX_STREAM."""
    if stream == CTX_STREAM:
        return None
    return stream if stream else CUDA_STREAM_LEGACY


class TutuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libtutu_b200 error {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    """Loads libtutu_b200.so; raises if it is not built (no fallback)."""
    global _lib
    if _lib is None:
        import os
        path = Path(os.environ.get("TUTU_LIB", LIB_PATH))  # TUTU_LIB: experiment builds only
        if not path.exists():
            raise RuntimeError(f"{path} is missing: build it with `python -m tuturenderer_b200.build` "
                               "(the CUDA library is the product; there is no CPU fallback)")
        l = C.CDLL(str(path))
        for name, (res, args) in ABI.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _check(rc: int, ctx=None) -> None:
    if rc != TUTU_OK:
        msg = lib().tutu_last_error(ctx)
        raise TutuError(rc, msg.decode() if msg else "?")


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


# ------------------------------------------------------------------------------------------------
# scene description (host side)
# ------------------------------------------------------------------------------------------------
@dataclass
class Scene:
    """Python mirror of TutuSceneDesc: numpy arrays in Scene::objList order."""
    prims: np.ndarray
    materials: np.ndarray
    eye: tuple = (0.0, 0.0, 0.0)
    viewdir: tuple = (0.0, 0.0, 1.0)
    updir: tuple = (0.0, 1.0, 0.0)
    hfov_deg: int = 40
    width: int = 64
    height: int = 64
    parallel_projection: int = 0
    bkgcolor: tuple = (0.0, 0.0, 0.0)
    eta: float = 1.0
    bvh_nodes: np.ndarray | None = None
    textures: list = field(default_factory=lambda: [[], [], [], []])  # per channel: list of (h,w,3) float32

    def __post_init__(self):
        self.prims = np.ascontiguousarray(self.prims, dtype=PRIM_DTYPE)
        self.materials = np.ascontiguousarray(self.materials, dtype=MATERIAL_DTYPE)
        if self.bvh_nodes is not None:
            self.bvh_nodes = np.ascontiguousarray(self.bvh_nodes, dtype=BVHNODE_DTYPE)

    def camera_struct(self) -> TutuCamera:
        cam = TutuCamera()
        cam.eye[:] = self.eye
        cam.viewdir[:] = self.viewdir
        cam.updir[:] = self.updir
        cam.hfov_deg = int(self.hfov_deg)
        cam.width, cam.height = int(self.width), int(self.height)
        cam.parallel_projection = int(self.parallel_projection)
        return cam

    def to_c(self):
        """Returns (TutuSceneDesc, keepalive)."""
        d = TutuSceneDesc()
        keep = [self.prims, self.materials]
        d.struct_size = C.sizeof(TutuSceneDesc)
        d.n_prims = len(self.prims)
        d.prims = _ptr(self.prims) if len(self.prims) else None
        d.n_materials = len(self.materials)
        d.materials = _ptr(self.materials) if len(self.materials) else None
        if self.bvh_nodes is not None and len(self.bvh_nodes):
            d.n_bvh_nodes = len(self.bvh_nodes)
            d.bvh_nodes = _ptr(self.bvh_nodes)
            keep.append(self.bvh_nodes)
        for c in range(4):
            texs = self.textures[c]
            d.n_tex[c] = len(texs)
            if texs:
                arr = (TutuTexture * len(texs))()
                for i, t in enumerate(texs):
                    t = np.ascontiguousarray(t, dtype=np.float32)
                    keep.append(t)
                    arr[i].height, arr[i].width = t.shape[0], t.shape[1]
                    arr[i].rgb = t.ctypes.data_as(C.POINTER(C.c_float))
                keep.append(arr)
                d.tex[c] = C.cast(arr, C.POINTER(TutuTexture))
        d.camera = self.camera_struct()
        d.bkgcolor[:] = self.bkgcolor
        d.eta = float(self.eta)
        return d, keep

    def save(self, path) -> None:
        d, _keep = self.to_c()
        _check(lib().tutu_scene_file_save(C.byref(d), str(path).encode()))

    @staticmethod
    def from_desc(d: TutuSceneDesc) -> "Scene":
        prims = np.ctypeslib.as_array(C.cast(d.prims, C.POINTER(C.c_uint8)), (d.n_prims * PRIM_DTYPE.itemsize,)).view(PRIM_DTYPE).copy() if d.n_prims else np.zeros(0, PRIM_DTYPE)
        mats = np.ctypeslib.as_array(C.cast(d.materials, C.POINTER(C.c_uint8)), (d.n_materials * MATERIAL_DTYPE.itemsize,)).view(MATERIAL_DTYPE).copy() if d.n_materials else np.zeros(0, MATERIAL_DTYPE)
        nodes = None
        if d.n_bvh_nodes and d.bvh_nodes:
            nodes = np.ctypeslib.as_array(C.cast(d.bvh_nodes, C.POINTER(C.c_uint8)), (d.n_bvh_nodes * 12,)).view(BVHNODE_DTYPE).copy()
        textures = [[], [], [], []]
        for c in range(4):
            for i in range(d.n_tex[c]):
                t = d.tex[c][i]
                n = t.width * t.height * 3
                a = np.ctypeslib.as_array(t.rgb, (n,)).copy().reshape(t.height, t.width, 3) if n else np.zeros((t.height, t.width, 3), np.float32)
                textures[c].append(a)
        cam = d.camera
        return Scene(prims=prims, materials=mats, eye=tuple(cam.eye), viewdir=tuple(cam.viewdir),
                     updir=tuple(cam.updir), hfov_deg=cam.hfov_deg, width=cam.width, height=cam.height,
                     parallel_projection=cam.parallel_projection, bkgcolor=tuple(d.bkgcolor), eta=d.eta,
                     bvh_nodes=nodes, textures=textures)

    @staticmethod
    def load(path) -> "Scene":
        h = _P()
        _check(lib().tutu_scene_file_load(str(path).encode(), C.byref(h)))
        try:
            return Scene.from_desc(lib().tutu_scene_file_desc(h).contents)
        finally:
            lib().tutu_scene_file_free(h)

    def with_size(self, width: int, height: int) -> "Scene":
        import copy
        s = copy.copy(self)
        s.width, s.height = width, height
        return s


def bvh_build(prims: np.ndarray) -> np.ndarray:
    """Midpoint BVH with the reference's split rule (host only)."""
    prims = np.ascontiguousarray(prims, dtype=PRIM_DTYPE)
    n = len(prims)
    out = np.zeros(max(2 * n - 1, 1), BVHNODE_DTYPE)
    cnt = C.c_uint32(0)
    _check(lib().tutu_bvh_build(_ptr(prims) if n else None, n, _ptr(out), C.byref(cnt)))
    return out[:cnt.value]


def traversal_tree_check(scene: "Scene") -> dict:
    """Host-only self check of the traversal trees an upload of `scene` builds (tutu_traversal_tree_check)."""
    d, keep = scene.to_c()
    out = TutuTreeCheck()
    _check(lib().tutu_traversal_tree_check(C.byref(d), C.byref(out)))
    del keep
    return {n: getattr(out, n) for n, _ in TutuTreeCheck._fields_}


def synth_heightfield(G: int, seed: int = 12345) -> np.ndarray:
    prims = np.zeros(2 * G * G, PRIM_DTYPE)
    _check(lib().tutu_synth_heightfield(G, seed, _ptr(prims)))
    return prims


def synth_rays(kind: int, n: int, seed: int = 12345, first: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    if out is None:
        out = np.empty((n, RAY_FLOATS), np.float32)
    _check(lib().tutu_synth_rays(kind, seed, first, n, _ptr(out)))
    return out


def guard_check() -> tuple[int, int]:
    """(live device allocations, overwritten guard-band bytes) of a -DTUTU_GUARDS build; TutuError otherwise."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    _check(lib().tutu_debug_guard_check(C.byref(a), C.byref(b)))
    return a.value, b.value


def write_ppm(path, rgb8: np.ndarray, binary: bool = False) -> None:
    """8-bit image (H, W, 3) -> PPM; binary=False is the reference's ASCII P3 byte for byte."""
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    h, w = rgb8.shape[:2]
    _check(lib().tutu_write_ppm(str(path).encode(), w, h, _ptr(rgb8), int(binary)))


def write_png(path, rgb8: np.ndarray) -> None:
    rgb8 = np.ascontiguousarray(rgb8, np.uint8)
    h, w = rgb8.shape[:2]
    _check(lib().tutu_write_png(str(path).encode(), w, h, _ptr(rgb8)))


def load_texture(path, normal_map: bool = False) -> np.ndarray:
    """Texture file (ASCII P3 as the reference reads it, binary P6, PNG) -> (H, W, 3) float32 texels for
    Scene.textures; normal_map applies the reference's `bump` recovery c * 2 - 1."""
    p = C.POINTER(C.c_float)()
    w, h = C.c_int32(0), C.c_int32(0)
    _check(lib().tutu_texture_load(str(path).encode(), int(normal_map), C.byref(p), C.byref(w), C.byref(h)))
    try:
        return np.ctypeslib.as_array(p, (h.value, w.value, 3)).copy()
    finally:
        lib().tutu_texture_free(p)


def post_params(**kw) -> TutuPostParams:
    """The reference's Postprocessor constants (Postprocessor.hpp:10-14), optionally overridden."""
    p = TutuPostParams()
    lib().tutu_post_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def default_material(**kw) -> np.ndarray:
    """One TutuMaterial with the reference's defaults (Material.hpp:21-30)."""
    m = np.zeros(1, MATERIAL_DTYPE)
    m["diffuse"] = (0.9, 0.9, 0.9)
    m["specular"] = (1.0, 1.0, 1.0)
    m["type"] = MAT_LAMBERTIAN
    m["alpha"], m["eta"], m["roughness"], m["metallic"] = 1.0, 1.0, 1.0, 0.0
    for k, v in kw.items():
        m[k] = v
    return m


# ------------------------------------------------------------------------------------------------
# device context
# ------------------------------------------------------------------------------------------------
class Context:
    """Owns a TutuCtx.  Mirrors the two reference plugin interfaces:

    * ``trace_closest`` / ``trace_any``  = IIntersectStrategy::UpdateInter / hasIntersection per batch
    * ``render_path``                    = IIntegrator::integrate of the PathTracing integrator
    """

    def __init__(self, device: int = 0):
        self._h = _P()
        _check(lib().tutu_ctx_create(device, C.byref(self._h)))
        self.device = device
        self.scene: Scene | None = None

    def close(self):
        if self._h:
            lib().tutu_ctx_destroy(self._h)
            self._h = _P()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _check(rc, self._h)

    # ---- scene
    def upload(self, scene: Scene) -> None:
        d, keep = scene.to_c()
        self._ck(lib().tutu_scene_upload(self._h, C.byref(d)))
        del keep
        self.scene = scene

    def builder(self, name: str = "auto") -> None:
        """Who builds the traversal tree at the next upload: 'auto', 'host_sah' or 'device_lbvh'."""
        self._ck(lib().tutu_scene_builder(self._h, BUILDERS[name]))

    def upload_stats(self) -> dict:
        u = TutuUploadStats()
        self._ck(lib().tutu_upload_stats(self._h, C.byref(u)))
        d = {n: getattr(u, n) for n, _ in TutuUploadStats._fields_}
        d["builder"] = {v: k for k, v in BUILDERS.items()}.get(d["builder"], d["builder"])
        return d

    def set_camera(self, scene: Scene) -> None:
        cam = scene.camera_struct()
        self._ck(lib().tutu_scene_set_camera(self._h, C.byref(cam)))

    def info(self) -> TutuSceneInfo:
        i = TutuSceneInfo()
        self._ck(lib().tutu_scene_info(self._h, C.byref(i)))
        return i

    def node_bytes(self) -> dict:
        """Record sizes of the tree the production walk descends (for the algorithmic-bytes figure)."""
        i = self.info()
        return {"node": int(i.trav_node_bytes), "leaf": int(i.trav_leaf_bytes)}

    def set_traversal_mode(self, mode: int) -> None:
        self._ck(lib().tutu_set_traversal_mode(self._h, mode))

    STACKS = {"auto": 0, "shared": 1, "local": 2}

    def traversal_stack(self, where: str = "auto") -> None:
        """Traversal stack of the tree kernels: 'auto' (by the size of the traversal arrays), 'shared', 'local'."""
        self._ck(lib().tutu_traversal_stack(self._h, self.STACKS[where]))

    TRACERS = {"auto": 0, "packets": 1, "lanes": 2}

    def bdpt_queue_tracer(self, name: str = "auto") -> None:
        """Queue tracers of render_bdpt on scenes with a tree: 'auto' (measured per scene), 'packets', 'lanes'."""
        self._ck(lib().tutu_bdpt_queue_tracer(self._h, self.TRACERS[name]))

    def bdpt_queue_tracer_measured(self) -> dict:
        t, a, b = C.c_int(0), C.c_float(0), C.c_float(0)
        self._ck(lib().tutu_bdpt_queue_tracer_measured(self._h, C.byref(t), C.byref(a), C.byref(b)))
        return {"tracer": {-1: None, 0: "packets", 1: "lanes"}[t.value], "ms_packets": a.value, "ms_lanes": b.value}

    # ---- ray batches (host buffers)
    def trace_closest(self, rays: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, RAY_FLOATS)
        if out is None:
            out = np.empty(len(rays), HIT_DTYPE)
        self._ck(lib().tutu_trace_closest(self._h, _ptr(rays), len(rays), _ptr(out)))
        return out

    def trace_any(self, rays: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, RAY_FLOATS)
        if out is None:
            out = np.empty(len(rays), np.uint8)
        self._ck(lib().tutu_trace_any(self._h, _ptr(rays), len(rays), _ptr(out)))
        return out

    # ---- ray batches (raw pointers: device or pinned host, as the entry point says)
    def trace_closest_ptr(self, rays_ptr: int, n: int, hits_ptr: int) -> None:
        self._ck(lib().tutu_trace_closest(self._h, rays_ptr, n, hits_ptr))

    def trace_any_ptr(self, rays_ptr: int, n: int, out_ptr: int) -> None:
        self._ck(lib().tutu_trace_any(self._h, rays_ptr, n, out_ptr))

    def trace_closest_device(self, d_rays: int, n: int, d_hits: int, stream: int = 0) -> None:
        self._ck(lib().tutu_trace_closest_device(self._h, d_rays, n, d_hits, stream_handle(stream)))

    def trace_any_device(self, d_rays: int, n: int, d_out: int, stream: int = 0) -> None:
        self._ck(lib().tutu_trace_any_device(self._h, d_rays, n, d_out, stream_handle(stream)))

    def count_visits(self, d_rays: int, n: int, any_hit: bool) -> tuple[int, int]:
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._ck(lib().tutu_trace_count_visits(self._h, d_rays, n, int(any_hit), C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- path tracing
    def configure(self, paths_in_flight: int = 0, profile_stages: bool = False, lanes: int = 0) -> None:
        self._ck(lib().tutu_render_configure(self._h, paths_in_flight, lanes, int(profile_stages)))

    PIPELINES = {"auto": 0, "wavefront": 1, "resident": 2}

    def pipeline(self, name: str = "auto") -> None:
        """Path-tracing pipeline: 'auto' (the resident kernel for renders below 768 Ki paths of scenes with <= 32
        primitives, else the wavefront), 'wavefront', or 'resident' (one persistent kernel with the path state
        in registers; scenes of <= 32 primitives only, renders fail on larger ones)."""
        self._ck(lib().tutu_render_pipeline(self._h, self.PIPELINES[name]))

    def render_path(self, spp: int, seed: int = 1, out: np.ndarray | None = None) -> np.ndarray:
        i = self.info()
        if out is None:
            out = np.empty((i.height, i.width, 3), np.float32)
        self._ck(lib().tutu_render_path(self._h, spp, seed, _ptr(out)))
        return out

    def render_path_ptr(self, spp: int, seed: int, out_ptr: int) -> None:
        self._ck(lib().tutu_render_path(self._h, spp, seed, out_ptr))

    def render_accumulate_device(self, sample_begin: int, sample_count: int, seed: int, d_accum: int,
                                 stream: int = 0) -> None:
        self._ck(lib().tutu_render_path_accumulate_device(self._h, sample_begin, sample_count, seed, d_accum,
                                                          stream_handle(stream)))

    def finalize_device(self, d_accum: int, inv_spp: float, d_out: int, stream: int = 0) -> None:
        self._ck(lib().tutu_finalize_device(self._h, d_accum, inv_spp, d_out, stream_handle(stream)))

    # ---- bidirectional path tracing (IIntegrator::integrate of the reference's BDPT)
    def render_bdpt(self, spp: int, seed: int = 1, out: np.ndarray | None = None) -> np.ndarray:
        i = self.info()
        if out is None:
            out = np.empty((i.height, i.width, 3), np.float32)
        self._ck(lib().tutu_render_bdpt(self._h, spp, seed, _ptr(out)))
        return out

    def render_bdpt_ptr(self, spp: int, seed: int, out_ptr: int) -> None:
        self._ck(lib().tutu_render_bdpt(self._h, spp, seed, out_ptr))

    def render_bdpt_accumulate_device(self, sample_begin: int, sample_count: int, seed: int, d_accum: int,
                                      stream: int = 0) -> None:
        self._ck(lib().tutu_render_bdpt_accumulate_device(self._h, sample_begin, sample_count, seed, d_accum,
                                                          stream_handle(stream)))

    def finalize_bdpt_device(self, d_accum: int, inv_spp: float, d_out: int, stream: int = 0) -> None:
        self._ck(lib().tutu_finalize_bdpt_device(self._h, d_accum, inv_spp, d_out, stream_handle(stream)))

    # ---- output stage (PPMGenerator::writePixel)
    def quantize(self, rgb: np.ndarray, gamma: float = 0.78) -> np.ndarray:
        rgb = np.ascontiguousarray(rgb, np.float32)
        out = np.empty(rgb.shape, np.uint8)
        self._ck(lib().tutu_quantize(self._h, _ptr(rgb), rgb.size // 3, gamma, _ptr(out)))
        return out

    def quantize_device(self, d_rgb: int, n_pixels: int, d_out: int, gamma: float = 0.78, stream: int = 0) -> None:
        self._ck(lib().tutu_quantize_device(self._h, d_rgb, n_pixels, gamma, d_out, stream_handle(stream)))

    # ---- output stage (Postprocessor: bloom / exposure tone map)
    def postprocess(self, rgb: np.ndarray, mode: str = "hdr_bloom", params: TutuPostParams | None = None) -> np.ndarray:
        """Host image (H, W, 3) through the reference's Postprocessor stages; params=None = its #define constants."""
        rgb = np.ascontiguousarray(rgb, np.float32)
        out = np.empty_like(rgb)
        self._ck(lib().tutu_postprocess(self._h, _ptr(rgb), rgb.shape[1], rgb.shape[0], POST_MODES[mode],
                                        C.byref(params) if params is not None else None, _ptr(out)))
        return out

    def postprocess_device(self, d_rgb: int, width: int, height: int, d_out: int, mode: str = "hdr_bloom",
                           params: TutuPostParams | None = None, stream: int = 0) -> None:
        self._ck(lib().tutu_postprocess_device(self._h, d_rgb, width, height, POST_MODES[mode],
                                               C.byref(params) if params is not None else None, d_out, stream_handle(stream)))

    def stats(self) -> dict:
        s = TutuRenderStats()
        self._ck(lib().tutu_render_stats(self._h, C.byref(s)))
        return s.as_dict()

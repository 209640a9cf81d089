import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def api():
    """The ctypes binding of libtutu_b200.so; builds the library in-tree when it is stale."""
    from tuturenderer_b200 import build
    build.build_library()
    from tuturenderer_b200 import api as _api
    _api.lib()
    return _api


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build(ref=False)
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


@pytest.fixture(scope="session")
def cornell(api):
    return api.Scene.load(GOLDEN / "cornell_256.tscene")


@pytest.fixture(scope="session")
def mixed(api):
    return api.Scene.load(GOLDEN / "mixed.tscene")


@pytest.fixture(scope="session")
def hf24(api):
    return api.Scene.load(GOLDEN / "hf24.tscene")


@pytest.fixture()
def ctx(api):
    """A device context.  No skip when CUDA is missing: GPU tests must fail loudly, the product
    has no CPU fallback."""
    c = api.Context(0)
    yield c
    c.close()


def load_rays(path):
    return np.fromfile(path, np.float32).reshape(-1, 8)


def random_soup(api, n_tris, n_spheres=0, seed=0, extent=10.0, dup=0):
    """Random triangle soup (+ spheres) with `dup` exact duplicates (equal-t ties)."""
    rng = np.random.default_rng(seed)
    prims = np.zeros(n_tris + n_spheres + dup, api.PRIM_DTYPE)
    prims["tex_diffuse"] = prims["tex_normal"] = prims["tex_roughness"] = prims["tex_metallic"] = -1
    c = rng.uniform(0, extent, (n_tris, 1, 3))
    v = (c + rng.normal(0, extent * 0.08, (n_tris, 3, 3))).astype(np.float32)
    prims["v"][:n_tris] = v.reshape(n_tris, 9)
    n = np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0])
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20)
    prims["n"][:n_tris] = np.repeat(n[:, None, :], 3, 1).reshape(n_tris, 9)
    prims["uv"][:n_tris] = rng.uniform(0, 1, (n_tris, 6))
    for k in range(n_spheres):
        p = prims[n_tris + k:n_tris + k + 1]
        p["type"] = api.PRIM_SPHERE
        p["v"][0, 0:3] = rng.uniform(0, extent, 3)
        p["v"][0, 3] = rng.uniform(0.2, 1.0)
    for k in range(dup):
        prims[n_tris + n_spheres + k] = prims[rng.integers(0, n_tris)]
    return prims


def random_rays(n, seed=0, extent=10.0, tmax=(0.5, 12.0)):
    rng = np.random.default_rng(seed)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-0.2 * extent, 1.2 * extent, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 4:7] = d
    # Sphere::intersect takes A = 1, so directions are re-normalised in float like normalized()
    d32 = rays[:, 4:7]
    mag = np.sqrt((d32[:, 0] * d32[:, 0] + d32[:, 1] * d32[:, 1]) + d32[:, 2] * d32[:, 2]).astype(np.float32)
    rays[:, 4:7] = d32 * (np.float32(1) / mag)[:, None]
    rays[:, 7] = rng.uniform(tmax[0], tmax[1], n)
    return rays


def assert_hits_equal(a, b):
    """Bit-exact equality of two TutuHit arrays (prim, t, u, v)."""
    a8, b8 = a.view(np.uint8).reshape(len(a), 16), b.view(np.uint8).reshape(len(b), 16)
    bad = np.nonzero((a8 != b8).any(1))[0]
    assert len(bad) == 0, f"{len(bad)} of {len(a)} hits differ; first: {a[bad[:3]]} vs {b[bad[:3]]}"

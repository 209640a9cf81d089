"""tuturenderer_b200 — B200-native path-tracing core for TutuRenderer.

The product is the CUDA library ``libtutu_b200.so`` behind the C ABI in ``include/tutu_b200.h``
(sources under ``tuturenderer_b200/csrc``).  This package only builds it (:mod:`.build`) and binds
it with ctypes (:mod:`.api`) for the tests and the bench; there is no CPU fallback.
"""
from .api import (Context, Scene, TutuError, lib, bvh_build, synth_heightfield, synth_rays, default_material,
                  PRIM_DTYPE, MATERIAL_DTYPE, BVHNODE_DTYPE, HIT_DTYPE, RAY_FLOATS)

__all__ = ["Context", "Scene", "TutuError", "lib", "bvh_build", "synth_heightfield", "synth_rays",
           "default_material", "PRIM_DTYPE", "MATERIAL_DTYPE", "BVHNODE_DTYPE", "HIT_DTYPE", "RAY_FLOATS"]

"""The C-ABI library loads on a CPU-only box and exports every symbol include/tutu_b200.h declares.
No compute entry point is called here."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "tutu_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tutu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(api):
    l = api.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(l, n), f"libtutu_b200.so does not export {n}"
    assert set(names) == set(api.ABI), "api.ABI and include/tutu_b200.h disagree"


def test_abi_version(api):
    assert api.lib().tutu_abi_version() == 2


def test_pod_sizes_match_header(api):
    assert C.sizeof(api.TutuCamera) == 52
    assert C.sizeof(api.TutuSceneInfo) == 64
    assert C.sizeof(api.TutuPostParams) == 24
    assert api.PRIM_DTYPE.itemsize == 124 and api.MATERIAL_DTYPE.itemsize == 56


def test_shipped_library_has_no_environment_knobs(api):
    """The experiment knobs (TUTU_PRUNE_*, TUTU_NO_*, ...) and the losing traversal flavours are compiled only with
    -DTUTU_EXPERIMENTS; the shipped .so must not read the environment at all (only TUTU_LIB, in api.py, picks a build)."""
    blob = (ROOT / "tuturenderer_b200" / "libtutu_b200.so").read_bytes()
    for knob in (b"TUTU_PRUNE_REL", b"TUTU_PRUNE_ABS", b"TUTU_NO_SMALL", b"TUTU_NO_FAST_TREE", b"TUTU_GRID_DIV", b"TUTU_LEAF_BATCH",
                 b"TUTU_HOST_SLOTS", b"TUTU_SHADE_BLOCK_RT", b"TUTU_NO_CLASS_SORT", b"TUTU_REFILL_MIN"):
        assert knob not in blob, knob
    for sym in (b"trace_refill", b"trace_persistent", b"traverse_warp", b"k_trace_variant"):
        assert sym not in blob, sym


def test_no_gpu_is_an_error_not_a_fallback(api):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    with pytest.raises(api.TutuError) as e:
        api.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    for p in (ROOT / "tuturenderer_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".cpp", ".hpp", ".h"}:
            assert "oracle" not in p.read_text().replace("oracle harness", ""), f"{p} mentions the oracle"

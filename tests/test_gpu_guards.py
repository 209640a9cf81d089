"""Writes outside a device allocation.  compute-sanitizer is closed on the GPU pool, so the library carries its own
detector: a -DTUTU_GUARDS build puts every device allocation between two 64 KB bands of a byte pattern;
tools/guard_case.py drives every kernel family and host loop through that build (tiny queues, odd sizes, every
traversal mode and tree builder) and reads the bands back."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
GUARD_LIB = ROOT / "tuturenderer_b200" / "libtutu_b200_guard.so"


def test_shipped_library_has_no_guard_bands():
    from tuturenderer_b200 import api
    with pytest.raises(api.TutuError) as e:
        api.guard_check()
    assert e.value.code == api.TUTU_E_STATE


@pytest.mark.gpu
def test_no_kernel_writes_outside_its_allocations():
    if not GUARD_LIB.exists():  # __graft_entry__.build() makes it; a box without it compiles it (nvcc is in the image)
        from tuturenderer_b200 import build
        build.build_variant("guard", ["TUTU_GUARDS"])
    res = subprocess.run([sys.executable, str(ROOT / "tools" / "guard_case.py")], env={**os.environ, "TUTU_LIB": str(GUARD_LIB)},
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    for c in out["checks"]:
        assert c["bad_bytes"] == 0, c
        assert c["buffers"] >= 8, c
    assert out["bad_bytes"] == 0
    assert out["bad_bytes_after_poke"] == 3, "the detector missed a deliberate 3-byte overrun"

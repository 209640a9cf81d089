"""Builds tuturenderer_b200/libtutu_b200.so in-tree with nvcc for sm_100a.

The library is plain CUDA C++ behind a C ABI (include/tutu_b200.h); it does not link torch.
nvcc cross-compiles without a GPU, so this also runs on the CPU-only build box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libtutu_b200.so"

SOURCES = [CSRC / "tutu_b200.cu", CSRC / "resident.cu", CSRC / "trace_kernels.cu", CSRC / "device_bvh.cu", CSRC / "host_scene.cpp", CSRC / "host_image.cpp", CSRC / "host_wide.cpp"]
HEADERS = [CSRC / "trace.cuh", CSRC / "shade.cuh", CSRC / "vertex.cuh", CSRC / "wavefront.cuh", CSRC / "resident.cuh", CSRC / "bdpt.cuh", CSRC / "postprocess.cuh", CSRC / "wide.cuh", CSRC / "wf_types.cuh", CSRC / "trace_kernels.hpp", CSRC / "device_bvh.hpp", CSRC / "tutu_internal.hpp",
           ROOT / "include" / "tutu_b200.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libtutu_b200.so cannot be built")


def nvcc_command(out: Path = LIB, extra: list[str] | None = None) -> list[str]:
    return [
        nvcc_path(), "-O3", "-std=c++17", "--threads", "0",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lineinfo",
        # host code must round like the reference's scalar build (no FMA contraction)
        "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-pthread",
        "-shared", "-o", str(out),
        *(extra or []),
        *map(str, SOURCES),
        "-lz",  # host_image.cpp: PNG inflate / deflate
    ]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS + [Path(__file__)])


def build_variant(name: str, defines: list[str]) -> Path:
    """Experiment builds (tools/): libtutu_b200_<name>.so with extra -D flags, picked up through TUTU_LIB."""
    out = PKG / f"libtutu_b200_{name}.so"
    res = subprocess.run(nvcc_command(out, [f"-D{d}" for d in defines]), capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(res.stdout + res.stderr)
    return out


def build_guard_library(force: bool = False) -> Path:
    """libtutu_b200_guard.so (-DTUTU_GUARDS): the same sources with 64 KB pattern bands around every device allocation
    and a registry behind tutu_debug_guard_check; loaded only by tests/test_gpu_guards.py through TUTU_LIB."""
    out = PKG / "libtutu_b200_guard.so"
    if not force and out.exists():
        t = out.stat().st_mtime
        if not any(p.stat().st_mtime > t for p in SOURCES + HEADERS + [Path(__file__)]):
            return out
    return build_variant("guard", ["TUTU_GUARDS"])


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = nvcc_command(extra=["-Xptxas", "-v"] if verbose else None)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libtutu_b200.so:\n" + " ".join(cmd))
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
    print(LIB)

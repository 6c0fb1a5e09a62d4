"""Builds the C-ABI CUDA library ``thr3ed_atom_b200/_lib/libr3d_b200.so`` in-tree with nvcc for
sm_100a (B200).  nvcc cross-compiles without a GPU, so this runs in the CPU-only build container;
the resulting ``.so`` is git-ignored but travels to the GPU box with the repo snapshot.

    python -m thr3ed_atom_b200.build [--force]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_DIR = PKG_DIR / "_lib"
LIB_PATH = LIB_DIR / "libr3d_b200.so"
STAMP = LIB_DIR / "libr3d_b200.stamp"

SOURCES = ["r3d_api.cu", "r3d_render.cu", "r3d_aux.cu", "r3d_comm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def _extra_flags():
    """Extra nvcc flags from $R3D_NVCC_FLAGS (tuning experiments, e.g. "-DR3D_FWD_BLOCKS=5")."""
    return os.environ.get("R3D_NVCC_FLAGS", "").split()


def _source_digest() -> str:
    h = hashlib.sha256()
    files = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(INCLUDE.glob("*.h")))
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
    return h.hexdigest()


def is_current() -> bool:
    return LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == _source_digest()


AB_LIB_PATH = LIB_DIR / "libr3d_b200_ab.so"
AB_STAMP = LIB_DIR / "libr3d_b200_ab.stamp"


def build_ab(verbose: bool = False) -> Path:
    """Measurement build with every A/B kernel variant (-DR3D_AB_VARIANTS): ``R3D_LIB_PATH=<this file>`` selects it.  The
    product library (``build()``) carries only the kernels its dispatch uses and refuses ``R3dRenderConfig.variant != 0``."""
    LIB_DIR.mkdir(exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-DR3D_AB_VARIANTS", *_extra_flags(), f"-I{INCLUDE}", f"-I{CSRC}", "-o", str(AB_LIB_PATH)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(CSRC / s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        sys.stderr.write(proc.stderr)
    AB_STAMP.write_text(_source_digest())
    return AB_LIB_PATH


def ab_is_current() -> bool:
    """Was the measurement build made from the current sources?"""
    return AB_LIB_PATH.exists() and AB_STAMP.exists() and AB_STAMP.read_text().strip() == _source_digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if the sources changed since the last build; returns its path."""
    if not force and is_current():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, *_extra_flags(), f"-I{INCLUDE}", f"-I{CSRC}", "-o", str(LIB_PATH)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(CSRC / s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{' '.join(cmd)}\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        sys.stderr.write(proc.stderr)
    STAMP.write_text(_source_digest())
    return LIB_PATH


if __name__ == "__main__":
    if "--ab" in sys.argv:
        print(build_ab(verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

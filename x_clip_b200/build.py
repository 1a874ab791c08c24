"""In-tree build of libxclip_b200.so (the C-ABI library) with nvcc for sm_100a.

`python -m x_clip_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles
without a GPU; the resulting .so is git-ignored but travels with the source tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
# XCLIP_BUILD_ELECT=1 builds the experimental elect.sync variant NEXT TO the default library
# (own object directory, libxclip_b200_elect.so); `XCLIP_LIB_VARIANT=elect` makes _lib.py load it.
_VARIANT = "elect" if os.environ.get("XCLIP_BUILD_ELECT") == "1" else ""
OBJ = PKG / "csrc" / ("build_" + _VARIANT if _VARIANT else "build")
LIB = PKG / ("libxclip_b200_" + _VARIANT + ".so" if _VARIANT else "libxclip_b200.so")

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _extra_flags() -> list:
    """Opt-in code-generation switches (see csrc/common.cuh); empty for the default build."""
    flags = []
    if os.environ.get("XCLIP_BUILD_ELECT") == "1":
        flags.append("-DXCLIP_USE_ELECT=1")
    return flags


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: x_clip_b200 has no prebuilt or fallback path")


def _newest_header_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "xclip_b200.h"]
    return max(h.stat().st_mtime for h in hdrs)


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    dep_m = max(src.stat().st_mtime, _newest_header_mtime())
    if obj.exists() and obj.stat().st_mtime > dep_m:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, *_extra_flags(), "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr:
        print(r.stderr)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    srcs = sorted(CSRC.glob("*.cu"))
    if not srcs:
        raise RuntimeError("no CUDA sources found")
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if LIB.exists() and all(LIB.stat().st_mtime > o.stat().st_mtime for o in objs):
        return LIB
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs),
           "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(lib)

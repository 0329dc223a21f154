"""Builds libcc_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m cold_compress_b200.build [--force]

The built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
SO = PKG / "libcc_b200.so"
NVCC_FLAGS = [*(["-DCC_TOKEN_MICROTRACE"] if __import__("os").environ.get("CC_TOKEN_MICROTRACE") else []), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not SO.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*")) + [PKG.parent / "include" / "cc_b200.h"])
    return SO.stat().st_mtime < newest


def build_extension(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return SO
    cmd = ["nvcc", *NVCC_FLAGS, "-o", str(SO), *map(str, sources())]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed building libcc_b200.so")
    (PKG / "build_log.txt").write_text(res.stdout + res.stderr)
    return SO


if __name__ == "__main__":
    print(build_extension(force="--force" in sys.argv, verbose=True))

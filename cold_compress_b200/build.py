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
# CC_NVCC_DEFS="-DCC_TOKEN_PRODUCER=0 ...": A/B builds (python -m cold_compress_b200.build --out=profiles/libcc_x.so)
NVCC_FLAGS = [*(["-DCC_TOKEN_MICROTRACE"] if __import__("os").environ.get("CC_TOKEN_MICROTRACE") else []),
              *__import__("os").environ.get("CC_NVCC_DEFS", "").split(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not SO.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*")) + [PKG.parent / "include" / "cc_b200.h"])
    return SO.stat().st_mtime < newest


def build_extension(force: bool = False, verbose: bool = False, out: Path | None = None) -> Path:
    """Every .cu is compiled to its own object file (in parallel, only when stale), then linked into the library."""
    out = Path(out) if out else SO
    if not force and out == SO and not needs_build():
        return SO
    from concurrent.futures import ThreadPoolExecutor
    objdir = PKG.parent / "build" / ("obj" + ("_" + out.stem if out != SO else ""))
    objdir.mkdir(parents=True, exist_ok=True)
    headers = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "cc_b200.h"]
    hdr_time = max(h.stat().st_mtime for h in headers)
    flags = [f for f in NVCC_FLAGS if f != "--shared"]
    flag_stamp = objdir / "flags.txt"
    flags_changed = not flag_stamp.exists() or flag_stamp.read_text() != " ".join(flags)

    def compile_one(src: Path):
        obj, log = objdir / (src.stem + ".o"), objdir / (src.stem + ".log")
        if not force and not flags_changed and obj.exists() and log.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_time):
            return 0, log.read_text(), obj
        res = subprocess.run(["nvcc", *flags, "-c", "-o", str(obj), str(src)], capture_output=True, text=True)
        text = res.stdout + res.stderr
        if res.returncode == 0:
            log.write_text(text)
        return res.returncode, text, obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, sources()))
    logs = "".join(t for _, t, _ in results)
    if verbose or any(rc for rc, _, _ in results):
        sys.stderr.write(logs)
    if any(rc for rc, _, _ in results):
        raise RuntimeError("nvcc failed building libcc_b200.so")
    flag_stamp.write_text(" ".join(flags))
    res = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC", "-o", str(out),
                          *[str(o) for _, _, o in results]], capture_output=True, text=True)
    if res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libcc_b200.so")
    if out == SO:
        (PKG / "build_log.txt").write_text(logs)
    return out


if __name__ == "__main__":
    out = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")), None)
    print(build_extension(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, out=out))

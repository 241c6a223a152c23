"""Build libgaussigan_sm100.so in-tree with nvcc for sm_100a (B200).

``python gaussigan_b200/build.py [-v] [--force]`` or ``__graft_entry__.build()`` (run the file, not
``-m``: importing the package loads the library it is about to build).  nvcc cross-compiles without a
GPU.  The library lands in ``gaussigan_b200/lib/`` (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIBNAME = "libgaussigan_sm100.so"
SOURCES = ["gg_capi.cu", "gg_project.cu", "gg_splat.cu", "gg_splat_fast.cu", "gg_splat_tma.cu", "gg_fused.cu", "gg_geometry.cu", "gg_host.cu", "gg_p2p.cu"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libgaussigan_sm100.so cannot be built")


def _sources() -> list[Path]:
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                   + [PKG.parent / "include" / "gaussigan.h"])
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def lib_path() -> Path:
    return LIBDIR / LIBNAME


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    out = lib_path()
    stamp = LIBDIR / (LIBNAME + ".sha256")
    fp = _fingerprint()
    if not force and out.exists() and stamp.exists() and stamp.read_text().strip() == fp:
        return out
    # one object per translation unit, compiled concurrently (the serial build is ~45 s), then one link
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src: Path):
        obj = objdir / (src.stem + ".o")
        cmd = [_nvcc(), *cflags, "-c", "-o", str(obj), str(src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, res

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        done = list(pool.map(compile_one, _sources()))
    for src, obj, res in done:
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name} ({res.returncode}):\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(f"== {src.name}\n{res.stderr}", file=sys.stderr)
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *[str(o) for _, o, _ in done]]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    stamp.write_text(fp)
    return out


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)

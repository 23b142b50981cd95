"""Builds libfeonet_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libfeonet_b200.so")
SOURCES = ["feo_host.cpp", "feo_tiles.cpp", "feo_patch_plan.cpp", "feo_lattice_plan.cpp", "feo_kernels.cu", "feo_tiled.cu", "feo_patch.cu", "feo_lattice.cu", "feo_dense_tc.cu", "feo_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3", "-shared", "--use_fast_math=false", "-x", "cu",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libfeonet_b200.so cannot be built (there is no CPU fallback)")


HEADERS = ["feo_internal.h", "feo_lattice.h", "feo_patch.h", "feo_lattice_gen.inc"]
OBJ_DIR = os.path.join(PKG_DIR, "_build")  # per-source objects (git-ignored, not shipped): only changed sources recompile


def _deps_mtime() -> float:
    deps = [os.path.join(CSRC, f) for f in HEADERS] + [os.path.join(ROOT, "include", "feonet_b200.h"), os.path.abspath(__file__)]
    return max(os.path.getmtime(d) for d in deps if os.path.exists(d))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "feonet_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """One `nvcc -c` per source (in parallel, objects cached under _build/), then one `nvcc -shared` link."""
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f not in ("--use_fast_math=false", "-shared")]
    base = [_nvcc(), *flags, "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        base += ["-Xptxas", "-v"]
    hdr_t = _deps_mtime()

    def compile_one(src: str):
        path, obj = os.path.join(CSRC, src), os.path.join(OBJ_DIR, src + ".o")
        if not force and not verbose and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_t):
            return obj, None
        res = subprocess.run(base + ["-c", path, "-o", obj], capture_output=True, text=True)
        return obj, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for obj, res in results:
        if res is None:
            continue
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed building " + os.path.basename(obj))
        if verbose:
            sys.stderr.write(res.stderr)
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"] + [o for o, _ in results] + ["-o", LIB_PATH]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libfeonet_b200.so")
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose="-v" in sys.argv))

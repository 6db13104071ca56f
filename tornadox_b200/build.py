"""Build ``libtornadox_b200.so`` in-tree with nvcc for sm_100a (no JIT cache; the .so travels with the tree).

    python -m tornadox_b200.build [--force] [--verbose]
"""

import concurrent.futures
import hashlib
import os
import pathlib
import shutil
import subprocess
import sys

PKG = pathlib.Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "lib"
LIB = OUT_DIR / "libtornadox_b200.so"
OBJ_DIR = OUT_DIR / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; tornadox_b200 needs the CUDA toolkit to build its kernels")
    return exe


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _stamp():
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "tornadox_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    (obj.with_suffix(".ptxas.txt")).write_text(res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    return obj


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ (in parallel) and link the shared library.  Returns its path."""
    OUT_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(exist_ok=True)
    stamp_file = OUT_DIR / "build.stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    nvcc = _nvcc()
    jobs = [(nvcc, src, OBJ_DIR / (src.stem + ".o"), verbose) for src in _sources()]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(_compile_one, jobs))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs), "-lcudart"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    stamp_file.write_text(stamp)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

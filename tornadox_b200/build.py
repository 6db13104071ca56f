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
    nvcc, src, obj, verbose, defines = args
    cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    (obj.with_suffix(".ptxas.txt")).write_text(res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    return obj


def build(force=False, verbose=False, defines=(), variant=""):
    """Compile every .cu under csrc/ (in parallel) and link the shared library.  Returns its path.

    ``variant`` / ``defines`` build an experimental copy ``libtornadox_b200_<variant>.so`` with extra -D flags
    (tuning experiments; select it at run time with the environment variable TDX_LIB)."""
    OUT_DIR.mkdir(exist_ok=True)
    lib = LIB if not variant else OUT_DIR / f"libtornadox_b200_{variant}.so"
    obj_dir = OBJ_DIR if not variant else OUT_DIR / f"obj_{variant}"
    obj_dir.mkdir(exist_ok=True)
    stamp_file = OUT_DIR / ("build.stamp" if not variant else f"build_{variant}.stamp")
    stamp = _stamp() + " ".join(defines)
    if not force and lib.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return lib
    nvcc = _nvcc()
    jobs = [(nvcc, src, obj_dir / (src.stem + ".o"), verbose, list(defines)) for src in _sources()]
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(_compile_one, jobs))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), *map(str, objs), "-lcudart"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    stamp_file.write_text(stamp)
    return lib


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    var = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--variant=")), "")
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=extra, variant=var)
    print(path)

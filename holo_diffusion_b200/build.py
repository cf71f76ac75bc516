"""In-tree build of libholo_b200.so (sm_100a only).

`python holo_diffusion_b200/build.py` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libholo_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v" if os.environ.get("HOLO_PTXAS_V") else "-O3"]
FLAGS += os.environ.get("HOLO_NVCC_FLAGS", "").split()   # tuning experiments (-DHOLO_CONV_LDW=32 ...); part of the stamp

def _hash(paths) -> str:
    import hashlib
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def _read(path: str) -> str:
    return open(path).read().strip() if os.path.exists(path) else ""


def build(force: bool = False, verbose: bool = False) -> str:
    """Staleness is decided by CONTENT (flags + source + every header), per object and for the library: a snapshot
    copied to another box loses mtimes, and a flag change must recompile every object (an mtime check would keep the
    old device code under a new stamp)."""
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "holo_b200.h"))
    stamp_path = LIB + ".stamp"
    digest = _hash([os.path.join(CSRC, f) for f in srcs] + hdrs)
    if not force and os.path.exists(LIB) and _read(stamp_path) == digest:
        return LIB
    jobs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        od = _hash([src] + hdrs)
        if force or not os.path.exists(obj) or _read(obj + ".stamp") != od:
            jobs.append((src, obj, od))

    def cc(job):
        src, obj, od = job
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose or os.environ.get("HOLO_PTXAS_V"):
            sys.stderr.write(r.stderr)
        open(obj + ".stamp", "w").write(od)   # only after a compile that used the current FLAGS
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(cc, jobs))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    open(stamp_path, "w").write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

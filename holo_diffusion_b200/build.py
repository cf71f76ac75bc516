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
# Opt-in builds of code that has not run on a GPU yet (the default build keeps the validated device code):
#   HOLO_BUILD_PDL=1           programmatic dependent launch (csrc/common.cuh), switched on at run time with HOLO_PDL=1
#   HOLO_BUILD_SPLIT_KV=1      split-KV fused attention (csrc/attn_flash.cu), used with HOLO_ATTN_KV_SPLIT=auto|<n>
#   HOLO_BUILD_EXPERIMENTAL=1  both
_exp = os.environ.get("HOLO_BUILD_EXPERIMENTAL") == "1"
if _exp or os.environ.get("HOLO_BUILD_PDL") == "1":
    FLAGS.append("-DHOLO_ENABLE_PDL")
if _exp or os.environ.get("HOLO_BUILD_SPLIT_KV") == "1":
    FLAGS.append("-DHOLO_ENABLE_SPLIT_KV")


def _newer(src: str, dst: str, deps) -> bool:
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(p) > t for p in [src, *deps])


def _source_hash(paths) -> str:
    import hashlib
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for p in sorted(paths):
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "holo_b200.h"))
    # content stamp: a snapshot copied to another box loses mtimes; never recompile an up-to-date library there
    stamp_path = LIB + ".stamp"
    digest = _source_hash([os.path.join(CSRC, f) for f in srcs] + hdrs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp_path) and open(stamp_path).read().strip() == digest:
        return LIB
    jobs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or _newer(src, obj, hdrs):
            jobs.append((src, obj))

    def cc(job):
        src, obj = job
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose or os.environ.get("HOLO_PTXAS_V"):
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(cc, jobs))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if force or jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    open(stamp_path, "w").write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

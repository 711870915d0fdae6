"""In-tree build of the sm_100a C-ABI library ``libp2l.so`` (nvcc cross-compiles without a GPU).

    python -m pix2latent_b200.build            # incremental
    python -m pix2latent_b200.build --force

The library is plain ``extern "C"`` (see include/p2l.h); Python binds it with ctypes
(pix2latent_b200/_lib.py). Nothing here falls back to another backend.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libp2l.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-I", os.path.join(os.path.dirname(HERE), "include"),
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, force, hdr_m):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    srcp = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) > max(os.path.getmtime(srcp), hdr_m)):
        return obj, False
    cmd = [NVCC] + FLAGS + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, True


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    # objects built with other flags are stale whatever their timestamps say
    stamp = os.path.join(OBJ, "flags.txt")
    sig = " ".join(FLAGS)
    if not os.path.exists(stamp) or open(stamp).read() != sig:
        force = True
    hdr_m = _headers_mtime()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, hdr_m), srcs))
    objs = [o for o, _ in res]
    rebuilt = any(c for _, c in res)
    with open(stamp, "w") as f:
        f.write(sig)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("[p2l] built", LIB)
    elif verbose:
        print("[p2l] up to date:", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)

"""Builds the C-ABI CUDA library (sm_100a only) in-tree with nvcc.

The library is the product: if it is missing or stale the host module raises - there is no
PyTorch/CPU fallback for the hot path.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libseunet_b200.so")
STAMP = os.path.join(HERE, ".libseunet_b200.stamp")
SOURCES = ["conv_tc.cu", "wgrad_tc.cu", "pointwise.cu", "pointwise2.cu", "pointwise3.cu", "window.cu", "backward.cu", "backward2.cu", "loss.cu", "postproc.cu", "plan.cu"]  # missing files are skipped
INCLUDE = os.path.join(HERE, "..", "include")

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
]


def _source_hash(extra):
    # every file the translation units can see: all of csrc/ (sources, .cuh headers, the #included .inc schedules) and
    # include/ - a stale library after editing e.g. plan_bwd.inc would be loaded silently otherwise
    h = hashlib.sha256()
    for d in (CSRC, INCLUDE):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cuh", ".inc", ".h")):
                h.update(f.encode())
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + extra).encode())
    return h.hexdigest()


def build(force=False, bf16=None, verbose=False):
    """Compile libseunet_b200.so if sources changed. Returns the library path."""
    if bf16 is None:
        bf16 = os.environ.get("SEUNET_ACT_BF16", "0") == "1"
    extra = ["-DSEUNET_ACT_BF16"] if bf16 else []
    if os.environ.get("SEUNET_GRAD_BF16", "0") == "1":   # experiment only: bf16 gradient planes miss the 1e-2 gradient tolerance
        extra.append("-DSEUNET_GRAD_BF16")
    want = _source_hash(extra)
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == want:
                return LIB
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    # one nvcc per translation unit, in parallel (the conv / wgrad templates dominate: ~50 s serial, ~20 s parallel), then link
    objdir = os.path.join(HERE, ".build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas=-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        res = subprocess.run(["nvcc"] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        return obj, res

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, srcs))
    for obj, res in results:
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr, file=sys.stderr)
    res = subprocess.run(["nvcc", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a"] +
                         [o for o, _ in results] + ["-o", LIB], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    with open(STAMP, "w") as fh:
        fh.write(want)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

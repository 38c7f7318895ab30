"""Builds libsprl_b200.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m sprl_b200.build [--force]

The library has no CPU fallback and no torch dependency; it is loaded through
ctypes by sprl_b200.capi and linked by the C++ worker under sprl_b200/host/.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsprl_b200.so")
SOURCES = ["env.cu", "search.cu", "engine.cu", "evalnet.cu"]

BASE_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall,-Wno-unknown-pragmas",
]
# bit-exact fp32/fp64 against the reference's x86-64 build: no FMA contraction,
# IEEE division and square root, denormals kept
EXACT_FLAGS = ["-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"]
NVCC_FLAGS = BASE_FLAGS + EXACT_FLAGS
# the evaluator network is a floating-point contraction (tolerance, not bit parity): FMA allowed
FLAGS_OF = {"evalnet.cu": BASE_FLAGS}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.exists(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "sprl_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [_nvcc(), *FLAGS_OF.get(os.path.basename(src), NVCC_FLAGS), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose or out.strip():
            print(out)
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))

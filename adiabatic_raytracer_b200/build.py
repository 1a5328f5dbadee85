"""Builds the CUDA library in-tree: adiabatic_raytracer_b200/libart_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libart_b200.so")
# tuning experiments: ART_DEFINES="-DART_TRACE_MINBLOCKS=3" ART_SO=/path/variant.so python -m ...build --force
EXTRA = os.environ.get("ART_DEFINES", "").split()
SO = os.environ.get("ART_SO", SO)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # the image exports CXX=/opt/gcc/bin/g++, a wrapper without OpenMP specs; pin the host compiler
    "-ccbin", "/usr/bin/g++",
    "-Xcompiler", "-fPIC", "-shared",
    # tables in the constant bank are read by __host__ __device__ functions whose host half never runs in this library
    "-diag-suppress", "20091",
]


def sources():
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".hpp"))]
    srcs.append(os.path.join(HERE, "..", "include", "art.h"))
    return srcs


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [NVCC] + FLAGS + EXTRA + ["-o", SO, os.path.join(CSRC, "art_api.cu"), "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libart_b200.so")
    if verbose:
        print(r.stderr)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", SO)

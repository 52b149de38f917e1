"""Builds mesoengine_b200/libmeso_b200.so (hand-written CUDA for sm_100a + the C ABI) with nvcc, in-tree.

-fmad=false / -ffp-contract=off: no fused multiply-add on either side of the ABI, so fp32/fp64 results are
bit-identical to the CPU oracle (DESIGN.md "determinism").  -lineinfo keeps ncu's source page usable.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmeso_b200.so")
SOURCES = ["meso_capi.cu", "k_voxelize.cu", "k_occupancy.cu", "k_raymarch.cu", "k_mesh.cu", "k_carve.cu", "k_resident.cu", "k_cubes.cu", "meso_group.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "--shared", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
    "-cudart", "static",
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "meso_cuda.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libmeso_b200.so")
    return OUT


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(OUT)

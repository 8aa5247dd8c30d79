"""Build libbcg_b200.so (hand-written sm_100a CUDA + the C-ABI of include/bcg_b200.h) in-tree.

    python -m bc_gym_planning_env_b200.csrc.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  --fmad=false keeps fp64 rounding operation-by-operation like the
NumPy reference (the kernels are memory-bound; the fused-multiply-add the reference's BLAS uses is
spelled fma() where it matters).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = [os.path.join(HERE, "bcg_kernels.cu")]
HEADERS = [os.path.join(HERE, "bcg_device.cuh"), os.path.join(HERE, "bcg_generate.cuh"),
           os.path.join(ROOT, "include", "bcg_b200.h")]
TARGET = os.path.join(HERE, "libbcg_b200.so")


def nvcc_command(verbose=False, target=None, defines=()):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false",
           "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", HERE,
           "-o", target or TARGET] + ["-D" + d for d in defines] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    return cmd


def needs_build():
    if not os.path.exists(TARGET):
        return True
    t = os.path.getmtime(TARGET)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return TARGET
    cmd = nvcc_command(verbose)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building %s" % TARGET)
    return TARGET


def build_variant(name, defines, verbose=False):
    """A tuning variant of the library (-D overrides of kernel shape macros) next to the product build:
    csrc/variants/libbcg_b200_<name>.so; load it with BCG_B200_LIB=<path> (profiles/probes/*)."""
    os.makedirs(os.path.join(HERE, "variants"), exist_ok=True)
    target = os.path.join(HERE, "variants", "libbcg_b200_%s.so" % name)
    cmd = nvcc_command(verbose, target, defines)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building %s" % target)
    return target


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(TARGET)

"""In-tree build recipes (nvcc cross-compiles sm_100a without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_DC = os.path.join(HERE, "libdeepcut_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build_kernels(force=False):
    csrc = os.path.join(HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc))] + [os.path.join(ROOT, "include", "deepcut_b200.h")]
    if not force and _newer(LIB_DC, srcs):
        return LIB_DC
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    _run([nvcc] + NVCC_FLAGS + ["-o", LIB_DC, os.path.join(csrc, "dc_abi.cu")])
    return LIB_DC


if __name__ == "__main__":
    print(build_kernels(force="--force" in sys.argv))

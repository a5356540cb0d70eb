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
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    extra = os.environ.get("DC_EXTRA_NVCC_FLAGS", "").split()        # A/B builds (e.g. -DDC_PTX_NO_CACHE_HINT)
    _run([nvcc] + NVCC_FLAGS + extra + ["-o", LIB_DC, os.path.join(csrc, "dc_abi.cu")])
    return LIB_DC


LIB_HOST = os.path.join(HERE, "libcaffe_b200.so")


def build_host(force=False):
    """C++ Caffe host (Net/Layer/Blob, planner, C binding): plain g++, links only the C ABI library."""
    hdir = os.path.join(HERE, "caffe_host")
    srcs = [os.path.join(hdir, "src", f) for f in sorted(os.listdir(os.path.join(hdir, "src"))) if f.endswith(".cpp")]
    deps = list(srcs)
    for root, _, files in os.walk(os.path.join(hdir, "include")):
        deps += [os.path.join(root, f) for f in files]
    deps += [os.path.join(ROOT, "include", "deepcut_b200.h"), os.path.join(ROOT, "include", "caffe_b200_c.h")]
    if not force and _newer(LIB_HOST, deps) and os.path.getmtime(LIB_HOST) >= os.path.getmtime(LIB_DC):
        return LIB_HOST
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"   # not $CXX: /opt/gcc links libstdc++ statically
    flags = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-I" + os.path.join(hdir, "include"), "-I" + os.path.join(ROOT, "include")]
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-4] + ".o")
        objs.append(o)
        if force or not _newer(o, deps):
            procs.append((s, subprocess.Popen([cxx] + flags + ["-c", s, "-o", o], stdout=subprocess.PIPE,
                                              stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("build failed: " + s)
    _run([cxx, "-shared", "-o", LIB_HOST] + objs + ["-L" + HERE, "-ldeepcut_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB_HOST


def build_all(force=False):
    build_kernels(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print(LIB_DC, LIB_HOST)

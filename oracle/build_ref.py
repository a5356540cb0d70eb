"""Builds oracle/_ref/librefcaffe.so: the reference's own CPU layer code compiled VERBATIM from where it lies
under /root/reference (never copied into this repo), with -DCPU_ONLY, against the stand-in headers of
oracle/ref_shim/ (glog, gflags, boost, cblas, and the protobuf-free caffe.pb.h look-alike) plus
oracle/ref_driver.cpp and proto_lite.cpp.  TEST INFRASTRUCTURE: used to validate the numpy restatement
(tests/test_oracle_ref.py) and as the `reference` CPU baseline of bench.py.  Outputs only into oracle/_ref/
(git-ignored; travels to the GPU box like any built .so).  Without /root/reference (the GPU box) it is a no-op
that keeps the prebuilt library."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "librefcaffe.so")

REF_SOURCES = ["src/caffe/common.cpp", "src/caffe/blob.cpp", "src/caffe/syncedmem.cpp", "src/caffe/layer.cpp",
               "src/caffe/util/math_functions.cpp", "src/caffe/util/im2col.cpp", "src/caffe/util/insert_splits.cpp"] + \
              ["src/caffe/layers/%s_layer.cpp" % n for n in ("base_conv", "conv", "deconv", "batch_norm", "scale", "bias", "relu", "eltwise",
                                                            "pooling", "crop", "sigmoid", "split", "neuron")]


def find_openblas():
    """OpenBLAS with plain cblas_* symbols inside the image's Python wheels (same path on the GPU box)."""
    import site
    for sp in site.getsitepackages() + [os.path.dirname(os.path.dirname(os.__file__)) + "/site-packages"]:
        for pat in ("opencv_python_headless.libs/libopenblas*.so*", "numpy.libs/libopenblas*.so*", "scipy.libs/libopenblas*.so*"):
            for p in sorted(glob.glob(os.path.join(sp, pat))):
                sym = subprocess.run(["nm", "-D", p], stdout=subprocess.PIPE, text=True).stdout
                if " cblas_sgemm\n" in sym or " T cblas_sgemm" in sym:
                    return p
    return None


def build(force=False):
    if not os.path.isdir(REF):
        return LIB if os.path.exists(LIB) else None
    os.makedirs(OUT, exist_ok=True)
    srcs = [os.path.join(REF, s) for s in REF_SOURCES] + [os.path.join(HERE, "ref_driver.cpp"),
                                                           os.path.join(ROOT, "deepcut-cnn_b200", "caffe_host", "src", "proto_lite.cpp")]
    deps = srcs + glob.glob(os.path.join(HERE, "ref_shim", "**", "*"), recursive=True)
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps if os.path.isfile(d)):
        return LIB
    blas = find_openblas()
    if blas is None:
        raise RuntimeError("no OpenBLAS with cblas_sgemm found in the Python wheels")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    flags = ["-std=c++14", "-O2", "-fPIC", "-fno-gnu-unique", "-DCPU_ONLY", "-w", "-I" + os.path.join(HERE, "ref_shim"), "-I" + os.path.join(REF, "include"),
             "-I" + os.path.join(ROOT, "deepcut-cnn_b200", "caffe_host", "include")]
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(OUT, os.path.basename(s)[:-4] + ".o")
        objs.append(o)
        procs.append((s, subprocess.Popen([cxx] + flags + ["-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            failed.append(s)
    if failed:
        raise RuntimeError("oracle/_ref: failed to compile " + ", ".join(failed))
    # only refcaffe_* is exported: the reference's caffe:: symbols must not meet the product's same-named ones when
    # both libraries sit in one test process
    vs = os.path.join(OUT, "exports.map")
    with open(vs, "w") as f:
        f.write("{ global: refcaffe_*; local: *; };\n")
    r = subprocess.run([cxx, "-shared", "-o", LIB, "-Wl,--version-script=" + vs, "-Wl,-Bsymbolic"] + objs + [blas, "-Wl,--disable-new-dtags", "-Wl,-rpath," + os.path.dirname(blas)], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("oracle/_ref: link failed")
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

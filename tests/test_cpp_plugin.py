"""The C++ plugin boundary from a user's side: tests/cpp/plugin_demo.cpp is compiled against this repo's Caffe headers,
defines and registers its own layer type with REGISTER_LAYER_CLASS, and drives caffe::Net through the reference's public C++ API.
Without a GPU: registry, Net::Init, shape propagation.  With one (-m gpu): Net::Forward in Caffe::GPU mode, Net::Reshape."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "plugin_demo.cpp")
PROTOTXT = os.path.join(ROOT, "tests", "cpp", "plugin_demo.prototxt")
LIBDIR = os.path.join(ROOT, "deepcut-cnn_b200")


def _build(tmp_path):
    exe = os.path.join(str(tmp_path), "plugin_demo")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"       # same compiler as the host library (build.py)
    cmd = [cxx, "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(LIBDIR, "caffe_host", "include"), "-I" + os.path.join(ROOT, "include"),
           SRC, "-o", exe, "-L" + LIBDIR, "-lcaffe_b200", "-ldeepcut_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    return exe


def _run(exe, mode):
    return subprocess.run([exe, mode, PROTOTXT], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)


def test_user_layer_registers_and_net_initialises(tmp_path):
    r = _run(_build(tmp_path), "init")
    assert r.returncode == 0 and "OK init: 4 layers" in r.stdout, r.stdout[-2000:]      # data's Split + ReLU + Sigmoid + ScaledSum


def test_cpu_mode_forward_fails_loudly(tmp_path):
    # the product has no CPU forward path: a user who forgets Caffe::set_mode(GPU) gets a CHECK failure (abort, like glog), not numbers
    src = os.path.join(str(tmp_path), "cpu_forward.cpp")
    open(src, "w").write('#include "caffe/caffe.hpp"\nint main(int, char** v) { caffe::Caffe::set_mode(caffe::Caffe::CPU); '
                         'caffe::Net<float> net(v[1], caffe::TEST); net.Forward(); return 0; }\n')
    exe = os.path.join(str(tmp_path), "cpu_forward")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([cxx, "-std=c++17", "-I" + os.path.join(LIBDIR, "caffe_host", "include"), "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                        "-L" + LIBDIR, "-lcaffe_b200", "-ldeepcut_b200", "-Wl,-rpath," + LIBDIR], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    tiny = os.path.join(str(tmp_path), "relu.prototxt")
    open(tiny, "w").write('input: "data"\ninput_dim: 1\ninput_dim: 1\ninput_dim: 2\ninput_dim: 2\nlayer { name: "r" type: "ReLU" bottom: "data" top: "r" }\n')
    r = subprocess.run([exe, tiny], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU forward path" in r.stdout, (r.returncode, r.stdout[-1500:])


@pytest.mark.gpu
def test_user_layer_runs_in_a_net_on_the_gpu(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = _run(_build(tmp_path), "gpu")
    assert r.returncode == 0 and "OK gpu" in r.stdout, r.stdout[-3000:]

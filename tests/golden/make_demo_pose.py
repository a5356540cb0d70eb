"""Generates tests/golden/demo_pose.npz by running the reference's demo UNCHANGED (python/pose/pose_demo.py under
/root/reference) in CPU mode against the product's `caffe` shim, with the forward computed by the reference's own CPU layers
(tests/demo_cpu/sitecustomize.py).  Also the helper module of tests/test_reference_demo_cpu.py.

    python tests/golden/make_demo_pose.py            # rewrites the fixture (needs /root/reference)

Fixture contents: `image` (uint8 HxWx3 RGB as stored in the PNG), `scales`, `pose` (5x14, the demo's output), `weights_seed`
(the calibrated synthetic weights are a pure function of the topology: synth.calibrated_weights).
"""
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden", "demo_pose.npz")


def synthetic_person(h=168, w=120, seed=12):
    """uint8 noise, the distribution the synthetic weights were calibrated on (synth.images): with random weights a smooth
    picture drives the stored BatchNorm statistics out of range and every sigmoid saturates."""
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def write_png(path, image):
    from PIL import Image
    Image.fromarray(np.ascontiguousarray(image), "RGB").save(path)


def prepare_workdir(base):
    """<base>/models/deepercut/ResNet-152.prototxt -> the reference's file; .caffemodel = calibrated synthetic weights written by
    the product (Net::ToProto); <base>/python/pose/ = the demo's working directory (it opens ../../models/...)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "deepcut-cnn_b200", "python"))
    import caffe
    synth = importlib.import_module("deepcut-cnn_b200.synth")
    ptx = importlib.import_module("deepcut-cnn_b200.prototxt")
    mdir = os.path.join(base, "models", "deepercut")
    os.makedirs(mdir, exist_ok=True)
    os.makedirs(os.path.join(base, "python", "pose"), exist_ok=True)
    proto = os.path.join(mdir, "ResNet-152.prototxt")
    if not os.path.exists(proto):
        os.symlink(os.path.join(REF, "models", "deepercut", "ResNet-152.prototxt"), proto)
    model = os.path.join(mdir, "ResNet-152.caffemodel")
    if not os.path.exists(model):
        caffe.set_mode_cpu()
        net = caffe.Net(proto, caffe.TEST)                      # the reference's own prototxt through the product's C++ parser
        net.set_params(synth.calibrated_weights(ptx.parse_file(proto)))
        net.save(model)
    return base


def run_demo(base, image_path, scales, visualize=True):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "demo_cpu"), os.path.join(ROOT, "deepcut-cnn_b200", "python"), ROOT])
    env["DC_TEST_CPU_FORWARD"] = "1"
    cmd = [sys.executable, os.path.join(REF, "python", "pose", "pose_demo.py"), image_path, "--scales", scales, "--use_cpu",
           "--visualize", "True" if visualize else "False"]
    r = subprocess.run(cmd, cwd=os.path.join(base, "python", "pose"), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-4000:]
    return np.load(image_path + "_pose.npz")["pose"]


if __name__ == "__main__":
    import tempfile
    sys.path.insert(0, ROOT)
    from oracle import build_ref
    build_ref.build()
    with tempfile.TemporaryDirectory() as tmp:
        base = prepare_workdir(tmp)
        image = synthetic_person()
        img = os.path.join(base, "python", "pose", "person.png")
        write_png(img, image)
        pose = run_demo(base, img, "1.,0.75")
        np.savez_compressed(OUT, image=image, scales=np.array([1.0, 0.75]), pose=pose)
        print("wrote", OUT, pose.shape, "\n", np.round(pose, 3))

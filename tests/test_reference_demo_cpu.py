"""The reference's demo, python/pose/pose_demo.py + estimate_pose.py, executed UNCHANGED (from /root/reference, where it lies)
on Python 3.12 against the product's `caffe` shim and its compat layer (deepcut-cnn_b200/python/compat).

There is no GPU in this container and no /root/reference on the GPU box, so the two halves meet through golden vectors:
HERE the demo runs in CPU mode (--use_cpu) with the forward computed by the reference's own CPU layers (tests/demo_cpu/
sitecustomize.py); everything else -- prototxt and .caffemodel loading, Blob reshape, data views, scipy.misc, the float
slicing of the tiled path -- is the product.  Its poses are the fixture tests/golden/demo_pose.npz (made by
tests/golden/make_demo_pose.py, which runs exactly this), and tests/test_demo_gpu.py holds the device pipeline to them.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
REF_DEMO = "/root/reference/python/pose/pose_demo.py"
GOLDEN = os.path.join(ROOT, "tests", "golden", "demo_pose.npz")

import make_demo_pose  # noqa: E402

pytestmark = pytest.mark.skipif(not os.path.exists(REF_DEMO), reason="the reference tree is not mounted")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    from oracle import build_ref, ref_caffe
    build_ref.build()
    if not ref_caffe.available():
        pytest.skip("oracle/_ref/librefcaffe.so not built")
    return make_demo_pose.prepare_workdir(str(tmp_path_factory.mktemp("demo")))


def test_demo_runs_unchanged_and_reproduces_the_golden_poses(workdir):
    g = np.load(GOLDEN)
    img = os.path.join(workdir, "python", "pose", "person.png")
    make_demo_pose.write_png(img, g["image"])
    pose = make_demo_pose.run_demo(workdir, img, "1.,0.75")
    assert pose.shape == (5, 14) and np.isfinite(pose).all()
    np.testing.assert_allclose(pose, g["pose"], rtol=0, atol=1e-4)
    assert os.path.exists(img + "_pose.npz_vis.png")          # scipy.misc.imsave through the compat layer


def test_tiled_path_runs_with_float_cutoffs(workdir):
    # > 700 px after the 64 px pad: _process_image_tiled splits the input and slices the tiles' maps with the FLOAT cut_off
    # (estimate_pose.py:167,251-255) -- a TypeError on NumPy 2 without the shim's loose-index arrays
    rng = np.random.default_rng(5)
    image = rng.integers(0, 256, (660, 200, 3), dtype=np.uint8)
    img = os.path.join(workdir, "python", "pose", "tall.png")
    make_demo_pose.write_png(img, image)
    pose = make_demo_pose.run_demo(workdir, img, "1.", visualize=False)
    assert pose.shape == (5, 14) and np.isfinite(pose).all()
    assert (pose[2] > 0).all() and (pose[2] < 1).all()

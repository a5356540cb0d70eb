"""The device pipeline (deepcut-cnn_b200/python/pose/estimate_pose.py: u8 upload -> dc_preprocess_u8_forward -> Net::Forward
-> dc_pose_from_maps) against the poses the REFERENCE'S DEMO ITSELF produced: tests/golden/demo_pose.npz was written by
/root/reference/python/pose/pose_demo.py run unchanged with the reference's CPU layers computing the forward
(tests/test_reference_demo_cpu.py, tests/golden/make_demo_pose.py).  Same image, same weights, scales 1.0 and 0.75."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dcutil
import netutil

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demo_pose.npz")


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def test_device_pipeline_reproduces_the_reference_demo_poses(tmp_path):
    g = np.load(GOLDEN)
    caffe = dcutil.caffe_module()
    from pose.estimate_pose import estimate_pose
    caffe.set_mode_gpu()
    caffe.set_device(0)
    path, weights = netutil.build(tmp_path, (3, 8, 36, 3), 688, 688)        # the shipped ResNet-152 deploy net
    image = g["image"][:, :, ::-1]                                            # RGB file -> BGR, pose_demo.py:116-121
    pose = estimate_pose(np.ascontiguousarray(image), path, None, [float(s) for s in g["scales"]], weights=weights)
    want = g["pose"]
    assert pose.shape == want.shape == (5, 14)
    print("\n[demo] max |dx,dy| %.4f px, max |dconf| %.2e" % (np.abs(pose[:2] - want[:2]).max(), np.abs(pose[2] - want[2]).max()))
    # the arg-max cell is the same for every joint (1e-3 on the maps cannot move a maximum that leads by more), so positions
    # differ only through loc_pred (x sqrt(53) x 1e-3) and confidences through prob
    assert np.abs(pose[2] - want[2]).max() < 1e-3
    assert np.abs(pose[:2] - want[:2]).max() < 2e-2
    assert np.abs(pose[3:] - want[3:]).max() < 2e-2

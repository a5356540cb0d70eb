"""Device pre-processing (dc_preprocess_u8_forward) and the device-side estimate_pose pipeline vs the CPU oracle.
Byte/integer work: BIT-EXACT against oracle/preprocess.py (itself pinned to Pillow) and the Pillow golden vectors."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dcutil
import gpuharness
import netutil
from dcutil import libdc
from oracle import caffe_ref, preprocess as pp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz")


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    gpuharness.init()


def run_preprocess(image, scale):
    L = libdc.lib()
    h, w = image.shape[:2]
    plan, oh, ow, ws = C.c_void_p(), C.c_int(), C.c_int(), C.c_size_t()
    libdc.check(L.dc_preprocess_plan_create(h, w, float(scale), C.byref(plan)))
    libdc.check(L.dc_preprocess_plan_info(plan, C.byref(oh), C.byref(ow), C.byref(ws)))
    d_img = gpuharness.dev(image)
    d_out = torch.full((3, oh.value, ow.value), float("nan"), dtype=torch.float32, device="cuda")
    d_ws = torch.empty(max(ws.value, 1), dtype=torch.uint8, device="cuda")
    mean = pp.MEAN.astype(np.float32)
    libdc.check(L.dc_preprocess_u8_forward(plan, C.c_void_p(d_img.data_ptr()), mean.ctypes.data_as(C.POINTER(C.c_float)),
                                           C.c_void_p(d_out.data_ptr()), C.c_void_p(d_ws.data_ptr()), gpuharness.stream_ptr()))
    torch.cuda.synchronize()
    libdc.check(L.dc_preprocess_plan_destroy(plan))
    return d_out.cpu().numpy()


@pytest.mark.parametrize("h,w,scale", [(45, 70, 1.0), (45, 70, 0.6), (45, 70, 1.45), (37, 53, 0.5), (64, 64, 2.0), (61, 35, 0.31), (1, 1, 1.0),
                                       (3, 2, 0.25), (360, 640, 0.85), (720, 1280, 1.0), (720, 1280, 0.5), (720, 1280, 1.5)])
def test_preprocess_bit_exact_vs_oracle(h, w, scale):
    image = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = pp.net_input_from_image(image, scale)
    got = run_preprocess(image, scale)
    assert got.shape == want.shape
    assert np.array_equal(got, want)


def test_preprocess_reproduces_pillow_golden_vectors():
    g = np.load(GOLDEN)
    for key in g.files:
        if key.startswith("scale_"):
            assert np.array_equal(run_preprocess(g["image"], float(key[6:])), g[key]), key


def test_preprocess_rejects_bad_arguments():
    L = libdc.lib()
    plan = C.c_void_p()
    assert L.dc_preprocess_plan_create(0, 10, 1.0, C.byref(plan)) != 0
    assert L.dc_preprocess_plan_create(10, 10, 0.0, C.byref(plan)) != 0
    assert L.dc_preprocess_plan_create(10, 10, 1e-4, C.byref(plan)) != 0        # rescales to zero pixels
    assert b"scale" in L.dc_last_error()


def test_estimate_pose_device_pipeline_matches_reference_pipeline(tmp_path):
    """uint8 image -> device pre-processing -> fused forward -> device read-out, against the reference's host pipeline
    restated on the oracle (oracle.preprocess -> CPU forward -> _pose_from_mats), over a two-scale pyramid."""
    path, weights = netutil.build(tmp_path, (1, 1, 1, 1), 64, 64)
    caffe = dcutil.caffe_module()          # puts deepcut-cnn_b200/python on sys.path (caffe, pose)
    caffe.set_mode_gpu()
    caffe.set_device(0)
    from pose import estimate_pose as ep
    image = np.random.default_rng(11).integers(0, 256, (90, 123, 3), dtype=np.uint8)
    scales = [1.0, 0.7]
    got = ep.estimate_pose(image, path, None, scales, weights=weights)
    best, best_conf = None, 0.0
    for s in scales:
        x = pp.net_input_from_image(image, s)[None]
        if netutil.reference_available():
            out = netutil.reference_forward(path, weights, x, want=["prob", "loc_pred"])
        else:
            out = netutil.oracle_forward(path, weights, x)
        pose = caffe_ref.pose_from_mats(out["prob"][0], out["loc_pred"][0], scale=s)
        if pose[2].min() > best_conf:
            best_conf, best = pose[2].min(), pose
    assert got.shape == (5, 14)
    # same arg-max cells (positions agree to the refinement's float error), confidences to the net's parity bound
    assert np.abs(got[2] - best[2]).max() < 1e-3
    assert np.abs(got[:2] - best[:2]).max() < 2e-2
    assert np.abs(got[3:] - best[3:]).max() < 2e-2

"""BASELINE configs[4], the scale pyramid over a batch (deepcut-cnn_b200/python/pose/pyramid.py): LPT-assigned (image, scale)
work items, one batched forward per geometry, best-of-scale per image -- must equal running the reference-signature
estimate_pose(image, ..., scales) image by image."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dcutil
import netutil


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def test_pyramid_batch_equals_per_image_estimate_pose(tmp_path):
    caffe = dcutil.caffe_module()
    from pose import pyramid
    from pose.estimate_pose import estimate_pose
    caffe.set_mode_gpu()
    caffe.set_device(0)
    path, weights = netutil.build(tmp_path, (1, 1, 1, 1), 64, 64)
    rng = np.random.default_rng(9)
    images = [rng.integers(0, 256, (96, 128, 3), dtype=np.uint8) for _ in range(3)] + [rng.integers(0, 256, (72, 88, 3), dtype=np.uint8)]
    scales = (0.5, 1.0, 1.5)
    best, poses, items = pyramid.estimate_poses_pyramid(images, path, None, scales=scales, weights=weights)
    assert len(best) == 4 and poses.shape == (12, 5, 14) and len(items) == 12
    # cost model: pixels of the rescaled, stride-aligned input (0.25 : 1 : 2.25 for one image size)
    c = [it[2] for it in items[:3]]
    assert c[0] < c[1] < c[2] and abs(c[2] / c[1] - 2.25) < 0.05
    for i, img in enumerate(images):
        want = estimate_pose(img, path, None, list(scales), weights=weights)
        assert best[i] is not None and want is not None
        np.testing.assert_allclose(best[i], want, rtol=0, atol=2e-3)
    # every item individually: the batched forward of 3 same-size images equals the single-image forwards
    for k, (i, s, _) in enumerate(items):
        one = estimate_pose(images[i], path, None, [s], weights=weights)
        if one is not None:
            np.testing.assert_allclose(poses[k], one, rtol=0, atol=2e-3)


def test_lpt_assignment_of_the_configs4_pyramid_is_balanced():
    from pose import pyramid
    import importlib
    dmod = importlib.import_module("deepcut-cnn_b200.dist")
    items = pyramid.work_items([(720, 1280)] * 8, (0.5, 1.0, 1.5))
    bins = dmod.lpt_assign([c for _, _, c in items], 8)
    loads = [sum(items[k][2] for k in b) for b in bins]
    assert max(loads) / min(loads) < 1.001 and all(len(b) == 3 for b in bins)

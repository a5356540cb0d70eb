"""Generates tests/golden/preprocess.npz with the REAL dependency chain of python/pose/estimate_pose.py:83-105 --
numpy padding + PIL.Image.resize(BILINEAR) (what scipy.misc.imresize called) + mean subtraction + crop -- for a seeded
uint8 image at three scales.  Needs Pillow (12.2.0 in this image).  Run: python tests/golden/make_golden_preprocess.py"""
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
MEAN = np.array([104., 117., 123.])


def reference_net_input(image, scale):
    h, w = image.shape[:2]
    bg_w = int(np.ceil(float(w) * scale / 8.) * 8.)
    bg_h = int(np.ceil(float(h) * scale / 8.) * 8.)
    img = np.vstack((image, np.tile(image[-1:, :, :], (64, 1, 1))))
    img = np.hstack((img, np.tile(img[:, -1:, :], (1, 64, 1))))
    size = tuple((np.array((img.shape[1], img.shape[0])) * scale).astype(int))          # scipy.misc.imresize, float size
    img = np.asarray(Image.fromarray(img).resize(size, resample=Image.BILINEAR)).astype('float32') - MEAN
    net_input = np.zeros((bg_h, bg_w, 3), dtype='float32')
    hh, ww = min(bg_h, img.shape[0]), min(bg_w, img.shape[1])
    net_input[:hh, :ww, :] = img[:hh, :ww, :]
    return net_input.transpose((2, 0, 1))


if __name__ == "__main__":
    image = np.random.default_rng(83).integers(0, 256, (45, 70, 3), dtype=np.uint8)
    out = {"image": image}
    for s in (1.0, 0.6, 1.45):
        out["scale_%g" % s] = reference_net_input(image, s)
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print({k: v.shape for k, v in out.items()})

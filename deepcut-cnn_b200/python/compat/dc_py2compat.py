"""Compatibility layer that lets the reference's demo -- python/pose/pose_demo.py and estimate_pose.py, Python-2 / SciPy-0.x
era code -- run UNCHANGED on this image's Python 3.12 / SciPy 1.18 / NumPy 2.3 against the B200 `caffe` shim.

Put this directory FIRST on PYTHONPATH (its sitecustomize.py calls install()), then the shim's python/ directory:

    cd <dir holding ../../models/deepercut/ResNet-152.{prototxt,caffemodel}>/python/pose
    PYTHONPATH=<repo>/deepcut-cnn_b200/python/compat:<repo>/deepcut-cnn_b200/python  python <reference>/python/pose/pose_demo.py image.png

What the demo needs that no longer exists:
  * ``scipy.misc.imread / imresize / imsave`` (pose_demo.py:123,149; estimate_pose.py:98) -- removed from SciPy 1.2/1.3.  They
    were thin wrappers over PIL (scipy/misc/pilutil.py of SciPy 0.19), restated here over Pillow: uint8 arrays go straight
    into ``Image.fromarray``; ``imresize(arr, fraction, interp='bilinear')`` is ``Image.resize((int(W*f), int(H*f)), BILINEAR)``.
  * float array indices (estimate_pose.py:167 ``cut_off = rf / stride`` is 28.0 -- ``_STRIDE = 8.`` -- and :251-255 slice with
    it): NumPy <= 1.11 truncated them with a DeprecationWarning.  The shim's ``Blob.data`` arrays (caffe/pycaffe.py
    _LooseIndexArray) accept integral floats as indices and stay that class through ``copy/transpose/concatenate``, which
    covers every array those lines touch.  Nothing is patched inside NumPy.
"""
import sys
import types

import numpy as np


def _pil():
    from PIL import Image
    return Image


def bytescale(data, cmin=None, cmax=None, high=255, low=0):
    """scipy.misc.bytescale (pilutil.py:35-100)."""
    data = np.asarray(data)
    if data.dtype == np.uint8:
        return data
    if cmin is None:
        cmin = data.min()
    if cmax is None:
        cmax = data.max()
    cscale = cmax - cmin
    if cscale == 0:
        cscale = 1
    scale = float(high - low) / cscale
    return ((data - cmin) * scale + low).clip(low, high).__add__(0.5).astype(np.uint8)


def toimage(arr, mode=None):
    """scipy.misc.toimage for the cases the demo produces: 2-D (greyscale) and HxWx3/4 arrays; non-uint8 data is byte-scaled."""
    Image = _pil()
    data = np.asarray(arr)
    if data.ndim == 2:
        if mode == "F":
            return Image.fromarray(data.astype(np.float32), "F")
        return Image.fromarray(bytescale(data), "L")
    if data.ndim == 3 and data.shape[2] in (3, 4):
        return Image.fromarray(np.ascontiguousarray(bytescale(data)), mode or ("RGB" if data.shape[2] == 3 else "RGBA"))
    raise ValueError("'arr' does not have a suitable array shape for any mode.")


def fromimage(im, flatten=False, mode=None):
    if mode is not None and mode != im.mode:
        im = im.convert(mode)
    elif im.mode == "P":
        im = im.convert("RGBA" if "transparency" in im.info else "RGB")
    if flatten:
        im = im.convert("F")
    elif im.mode == "1":
        im = im.convert("L")
    return np.array(im)


def imread(name, flatten=False, mode=None):
    """scipy.misc.imread (pilutil.py:103-156)."""
    im = _pil().open(name)
    return fromimage(im, flatten=flatten, mode=mode)


def imsave(name, arr, format=None):
    """scipy.misc.imsave (pilutil.py:159-200)."""
    im = toimage(arr, mode=None)
    if format is None:
        im.save(name)
    else:
        im.save(name, format)


_INTERP = {"nearest": 0, "lanczos": 1, "bilinear": 2, "bicubic": 3, "cubic": 3}


def imresize(arr, size, interp="bilinear", mode=None):
    """scipy.misc.imresize (pilutil.py:480-550): int = percent, float = fraction, tuple = (rows, cols)."""
    im = toimage(arr, mode=mode)
    if isinstance(size, (int, np.integer)):
        size = tuple((np.array(im.size) * (size / 100.0)).astype(int))
    elif isinstance(size, (float, np.floating)):
        size = tuple((np.array(im.size) * size).astype(int))
    else:
        size = (size[1], size[0])
    return fromimage(im.resize(tuple(int(s) for s in size), resample=_INTERP[interp]))


def install():
    """Publishes the functions above as ``scipy.misc`` (attribute + sys.modules entry).  Idempotent."""
    import scipy
    if getattr(getattr(scipy, "misc", None), "_dc_compat", False):
        return
    mod = types.ModuleType("scipy.misc", "scipy.misc image helpers restated over Pillow (deepcut-cnn_b200 compat layer)")
    for fn in (imread, imsave, imresize, toimage, fromimage, bytescale):
        setattr(mod, fn.__name__, fn)
    mod._dc_compat = True
    sys.modules["scipy.misc"] = mod
    scipy.misc = mod

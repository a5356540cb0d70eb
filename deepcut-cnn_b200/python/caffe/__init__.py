"""pycaffe-compatible `caffe` module over the B200 C++ host (libcaffe_b200.so).

Mirrors the names the DeeperCut demo uses from the reference's python/caffe/__init__.py:1-8 and
pycaffe.py: Net, TEST/TRAIN, set_mode_cpu/gpu, set_device.  Put this directory's parent on
PYTHONPATH (``deepcut-cnn_b200/python``) and ``import caffe``.
"""
from .pycaffe import Net, Blob, Layer
from ._caffe import (set_mode_cpu, set_mode_gpu, set_device, device_count, set_log_level, sync, TRAIN, TEST,
                     CaffeError, insert_splits_text)

__version__ = "1.0.0-b200"
__all__ = ["Net", "Blob", "Layer", "set_mode_cpu", "set_mode_gpu", "set_device", "device_count", "set_log_level",
           "sync", "TRAIN", "TEST", "CaffeError", "insert_splits_text"]

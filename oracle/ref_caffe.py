"""ctypes face of oracle/_ref/librefcaffe.so -- the reference's OWN CPU layer code (compiled verbatim from
/root/reference by oracle/build_ref.py) behind the same interface as the numpy restatement's ``RefNet``.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs; the product never loads it.  On the GPU box /root/reference does not exist; the
prebuilt library travels with the snapshot and this module only dlopen()s it.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "librefcaffe.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/librefcaffe.so is not built (python oracle/build_ref.py needs /root/reference)")
        l = ctypes.CDLL(LIB_PATH)
        vp, cp, ci, fp = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float)
        sigs = {
            "refcaffe_last_error": (cp, []), "refcaffe_net_create": (vp, [cp]), "refcaffe_net_destroy": (None, [vp]),
            "refcaffe_net_forward": (ci, [vp]), "refcaffe_num_layers": (ci, [vp]), "refcaffe_layer_name": (cp, [vp, ci]),
            "refcaffe_layer_type": (cp, [vp, ci]), "refcaffe_layer_num_blobs": (ci, [vp, ci]),
            "refcaffe_layer_blob_count": (ci, [vp, ci, ci]), "refcaffe_layer_blob_data": (fp, [vp, ci, ci]),
            "refcaffe_num_blobs": (ci, [vp]), "refcaffe_blob_name": (cp, [vp, ci]),
            "refcaffe_blob_shape": (ci, [vp, ci, ctypes.POINTER(ctypes.c_int)]),
            "refcaffe_blob_reshape": (ci, [vp, ci, ci, ci, ci, ci]), "refcaffe_blob_data": (fp, [vp, ci]),
        }
        for name, (res, args) in sigs.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


def _err():
    return lib().refcaffe_last_error().decode()


class RefCaffeNet:
    """One reference net (float, CPU, TEST phase).  ``forward`` mirrors oracle.caffe_ref.RefNet.forward."""

    def __init__(self, prototxt_text):
        self._h = lib().refcaffe_net_create(prototxt_text.encode())
        if not self._h:
            raise RuntimeError("reference net construction failed: " + _err())
        L = lib()
        self.layer_names = [L.refcaffe_layer_name(self._h, i).decode() for i in range(L.refcaffe_num_layers(self._h))]
        self.layer_types = [L.refcaffe_layer_type(self._h, i).decode() for i in range(len(self.layer_names))]
        self.blob_names = [L.refcaffe_blob_name(self._h, i).decode() for i in range(L.refcaffe_num_blobs(self._h))]
        self._blob_index = {n: i for i, n in enumerate(self.blob_names)}

    def __del__(self):
        if getattr(self, "_h", None):
            lib().refcaffe_net_destroy(self._h)
            self._h = None

    def _blob(self, name):
        i = self._blob_index[name]
        dims = (ctypes.c_int * 8)()
        n = lib().refcaffe_blob_shape(self._h, i, dims)
        shape = tuple(dims[a] for a in range(n))
        p = lib().refcaffe_blob_data(self._h, i)
        return np.ctypeslib.as_array(p, shape=(int(np.prod(shape)),)).reshape(shape)

    def set_params(self, weights):
        """weights: {layer name: [ndarray, ...]} in the reference's blob order."""
        L = lib()
        for li, name in enumerate(self.layer_names):
            if name not in weights:
                continue
            arrs = weights[name]
            assert len(arrs) == L.refcaffe_layer_num_blobs(self._h, li), name
            for j, a in enumerate(arrs):
                cnt = L.refcaffe_layer_blob_count(self._h, li, j)
                a = np.ascontiguousarray(a, np.float32).ravel()
                assert a.size == cnt, (name, j, a.size, cnt)
                dst = np.ctypeslib.as_array(L.refcaffe_layer_blob_data(self._h, li, j), shape=(cnt,))
                dst[:] = a

    def param_counts(self):
        L = lib()
        return {n: [L.refcaffe_layer_blob_count(self._h, li, j) for j in range(L.refcaffe_layer_num_blobs(self._h, li))]
                for li, n in enumerate(self.layer_names) if L.refcaffe_layer_num_blobs(self._h, li)}

    def forward(self, inputs, want=None):
        """inputs {name: NCHW ndarray}.  Returns {blob name: copy} for ``want`` (default: every blob whose name
        is not an auto-inserted split).  In-place layers overwrite their blob, so an in-place chain's name reads
        as its LAST value -- the same convention as pycaffe's net.blobs[...]."""
        for name, x in inputs.items():
            x = np.ascontiguousarray(x, np.float32)
            if lib().refcaffe_blob_reshape(self._h, self._blob_index[name], *x.shape):
                raise RuntimeError(_err())
        # Layer::Forward calls Reshape first (layer.hpp:456), so new input shapes propagate during the forward.
        for name, x in inputs.items():
            self._blob(name)[...] = x
        if lib().refcaffe_net_forward(self._h):
            raise RuntimeError("reference forward failed: " + _err())
        names = want if want is not None else [n for n in self.blob_names if "_split_" not in n]
        return {n: self._blob(n).copy() for n in names}

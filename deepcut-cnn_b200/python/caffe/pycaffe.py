"""Net / Blob / Layer with the pycaffe surface (reference python/caffe/pycaffe.py:22-108,
_caffe.cpp:225-276): OrderedDict views `blobs` and `params`, `inputs`/`outputs`, `forward(**kwargs)`,
`Blob.data` as a zero-copy NumPy view of mutable_cpu_data() that keeps the blob alive."""
import ctypes as C
import weakref
from collections import OrderedDict

import numpy as np

from . import _caffe
from ._caffe import lib, check, check_ptr, CaffeError


class Blob(object):
    def __init__(self, handle, net=None, index=-1):
        self._h = handle
        self._net = weakref.ref(net) if net is not None else None      # activation blobs know their Net (weakly: the blob
        self._index = index                                           # memory may outlive it, test_net.py:48-60)

    def __del__(self):
        try:
            lib.caffe_blob_release(self._h)
        except Exception:
            pass

    @property
    def shape(self):
        return tuple(lib.caffe_blob_shape(self._h, i) for i in range(lib.caffe_blob_num_axes(self._h)))

    @property
    def count(self):
        return lib.caffe_blob_count(self._h)

    def _legacy(self, i):
        s = self.shape
        return s[i] if i < len(s) else 1

    num = property(lambda self: self._legacy(0))
    channels = property(lambda self: self._legacy(1))
    height = property(lambda self: self._legacy(2))
    width = property(lambda self: self._legacy(3))

    def reshape(self, *dims):
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        arr = (C.c_int * len(dims))(*[int(d) for d in dims])
        check(lib.caffe_blob_reshape(self._h, len(dims), arr))

    def _view(self, ptr):
        shape = self.shape
        n = int(np.prod(shape)) if shape else 0
        if n == 0:
            return np.zeros(shape, np.float32)
        buf = (C.c_float * n).from_address(check_ptr(ptr))
        a = np.frombuffer(buf, dtype=np.float32).reshape(shape)
        a.flags.writeable = True
        self._keep = buf
        # the array holds a reference to this Blob object, which holds the shared_ptr handle
        return _BlobArray(a, self)

    def _check_fresh(self):
        net = self._net() if self._net is not None else None
        if net is not None and not lib.caffe_net_blob_fresh(net._h, self._index):
            raise CaffeError("blob '%s' was not written by the last forward: the fused B200 plan materialises only the net's outputs. "
                             "Ask for it -- net.forward(blobs=['%s']) -- or call net.materialize_intermediates(True) / "
                             "net.set_fusion(False) first." % (net._blob_names[self._index], net._blob_names[self._index]))

    @property
    def data(self):
        self._check_fresh()
        return self._view(lib.caffe_blob_mutable_cpu_data(self._h))

    @property
    def diff(self):
        return self._view(lib.caffe_blob_mutable_cpu_diff(self._h))

    def gpu_data_ptr(self):
        return check_ptr(lib.caffe_blob_gpu_data(self._h))

    def mutable_gpu_data_ptr(self):
        """Blob::mutable_gpu_data(): the device copy becomes the authoritative one (syncedmem.cpp:125-129)."""
        return check_ptr(lib.caffe_blob_mutable_gpu_data(self._h))

    def overwrite_gpu_data_ptr(self):
        """Device pointer for a writer of EVERY element (SyncedMemory::overwrite_gpu_data): no upload of a host copy first."""
        return check_ptr(lib.caffe_blob_overwrite_gpu_data(self._h))

    def cpu_data_ptr(self):
        """Blob::cpu_data() (const): syncs a device-side head to the host, does NOT count as a host write."""
        return check_ptr(lib.caffe_blob_cpu_data(self._h))

    HEADS = ("UNINITIALIZED", "HEAD_AT_CPU", "HEAD_AT_GPU", "SYNCED")

    @property
    def data_head(self):
        """SyncedMemory::head() of the data (include/caffe/syncedmem.hpp:65-66) as its enumerator's name."""
        h = lib.caffe_blob_data_head(self._h)
        if h < 0:
            raise CaffeError(lib.caffe_last_error().decode(errors="replace"))
        return self.HEADS[h]


def _loose(i):
    """NumPy <= 1.11 index semantics the reference demo relies on (estimate_pose.py:167,251-255: cut_off = rf / stride is the
    FLOAT 28.0 and is used as a slice bound): integral floats index like their int."""
    if isinstance(i, (float, np.floating)) and float(i).is_integer():
        return int(i)
    if isinstance(i, slice):
        return slice(_loose(i.start), _loose(i.stop), _loose(i.step))
    if isinstance(i, tuple):
        return tuple(_loose(j) for j in i)
    return i


class _LooseIndexArray(np.ndarray):
    """ndarray that accepts integral floats as indices and keeps doing so through copy / transpose / np.concatenate ..., i.e.
    for every array the demo derives from a ``Blob.data`` view."""
    def __getitem__(self, idx):
        return np.ndarray.__getitem__(self, _loose(idx))

    def __setitem__(self, idx, value):
        np.ndarray.__setitem__(self, _loose(idx), value)

    def __array_function__(self, func, types, args, kwargs):
        res = super().__array_function__(func, types, args, kwargs)
        if type(res) is np.ndarray:
            res = res.view(_LooseIndexArray)
        return res


class _BlobArray(_LooseIndexArray):
    """ndarray view that keeps its Blob (hence the C++ shared_ptr) alive."""
    def __new__(cls, arr, blob):
        obj = arr.view(cls)
        obj._blob = blob
        return obj

    def __array_finalize__(self, obj):
        self._blob = getattr(obj, "_blob", None)


class _Outputs(dict):
    """{blob name: array} whose arrays are taken from Blob.data on first access.  pycaffe returns
    eager `.data` views (pycaffe.py:106-108), which forces a device->host copy of EVERY output each
    forward -- 335 MB for the 364-channel next_pred the DeeperCut demo never reads
    (estimate_pose.py:231).  Same mapping, fetched lazily."""
    def __init__(self, net, names):
        dict.__init__(self, ((n, None) for n in names))
        self._net = net

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        if v is None:
            v = self._net.blobs[k].data
            dict.__setitem__(self, k, v)
        return v

    def get(self, k, default=None):
        return self[k] if k in self else default

    def items(self):
        return [(k, self[k]) for k in self]

    def values(self):
        return [self[k] for k in self]


class Layer(object):
    def __init__(self, net, index):
        self._net = net
        self._i = index
        self.type = lib.caffe_net_layer_type(net._h, index).decode()
        self.blobs = [Blob(check_ptr(lib.caffe_net_layer_blob(net._h, index, j)))
                      for j in range(lib.caffe_net_layer_num_blobs(net._h, index))]


class Net(object):
    def __init__(self, network_file, *args, **kwargs):
        """Net(network_file, phase) | Net(network_file, weights_file, phase) -- the two ctor forms
        of the reference binding (_caffe.cpp:76-96)."""
        weights = kwargs.get("weights")
        phase = kwargs.get("phase")
        if len(args) == 1:
            phase = args[0]
        elif len(args) == 2:
            weights, phase = args
        if phase is None:
            raise TypeError("Net(network_file, [weights_file,] phase)")
        with open(network_file):          # same failure mode as CheckFile (_caffe.cpp:45-52)
            pass
        self._network_file = network_file
        self._h = check_ptr(lib.caffe_net_create(network_file.encode(), int(phase)))
        if weights is not None:
            with open(weights):
                pass
            check(lib.caffe_net_copy_trained_from(self._h, weights.encode()))
        self._build_tables()

    @classmethod
    def from_string(cls, prototxt_text, phase):
        self = cls.__new__(cls)
        self._h = check_ptr(lib.caffe_net_create_from_string(prototxt_text.encode(), int(phase)))
        self._build_tables()
        return self

    def _build_tables(self):
        h = self._h
        self._blob_names = [lib.caffe_net_blob_name(h, i).decode() for i in range(lib.caffe_net_num_blobs(h))]
        self._layer_names = [lib.caffe_net_layer_name(h, i).decode() for i in range(lib.caffe_net_num_layers(h))]
        self._blobs = [Blob(check_ptr(lib.caffe_net_blob(h, i)), self, i) for i in range(len(self._blob_names))]
        self.layers = [Layer(self, i) for i in range(len(self._layer_names))]
        self._inputs = [lib.caffe_net_input_index(h, i) for i in range(lib.caffe_net_num_inputs(h))]
        self._outputs = [lib.caffe_net_output_index(h, i) for i in range(lib.caffe_net_num_outputs(h))]

    def __del__(self):
        try:
            lib.caffe_net_destroy(self._h)
        except Exception:
            pass

    # ---- pycaffe views (pycaffe.py:22-60)
    @property
    def blobs(self):
        return OrderedDict(zip(self._blob_names, self._blobs))

    @property
    def params(self):
        return OrderedDict((name, lr.blobs) for name, lr in zip(self._layer_names, self.layers) if len(lr.blobs) > 0)

    @property
    def inputs(self):
        return [self._blob_names[i] for i in self._inputs]

    @property
    def outputs(self):
        return [self._blob_names[i] for i in self._outputs]

    @property
    def name(self):
        return lib.caffe_net_name(self._h).decode()

    def bottom_names(self, layer_index):
        h = self._h
        return [self._blob_names[lib.caffe_net_layer_bottom_id(h, layer_index, j)]
                for j in range(lib.caffe_net_layer_num_bottoms(h, layer_index))]

    def top_names(self, layer_index):
        h = self._h
        return [self._blob_names[lib.caffe_net_layer_top_id(h, layer_index, j)]
                for j in range(lib.caffe_net_layer_num_tops(h, layer_index))]

    # ---- execution
    def _forward(self, start, end):
        check(lib.caffe_net_forward_from_to(self._h, start, end))

    def forward(self, blobs=None, start=None, end=None, **kwargs):
        """pycaffe.py:62-108: copy kwargs into input blobs, run, return {output name: array}."""
        if blobs is None:
            blobs = []
        if kwargs:
            if set(kwargs.keys()) != set(self.inputs):
                raise Exception("Input blob arguments do not match net inputs.")
            for in_, blob in kwargs.items():
                if blob.shape[0] != self.blobs[in_].num:
                    raise Exception("Input is not batch sized")
                self.blobs[in_].data[...] = blob
        if start is None and end is None:
            if set(blobs) - set(self.outputs) - set(self.inputs):
                # the caller wants intermediates (pycaffe.py:62-108 returns any named blob): the reference fills every blob on
                # every forward, the fused plan only the outputs -- run this call layer by layer, which materialises everything
                self._forward(0, len(self.layers) - 1)
            else:
                check(lib.caffe_net_forward(self._h))
            outputs = set(self.outputs + blobs)
        else:
            start_ind = 0 if start is None else self._layer_names.index(start)
            if end is None:
                end_ind = len(self.layers) - 1
                outputs = set(self.outputs + blobs)
            else:
                end_ind = self._layer_names.index(end)
                outputs = set([end] + blobs)
            self._forward(start_ind, end_ind)
        if self.fused_last_forward:
            outputs -= getattr(self, "_skipped", set())          # not computed (skip_outputs); per-layer forwards fill everything
        return _Outputs(self, sorted(outputs))

    def reshape(self):
        check(lib.caffe_net_reshape(self._h))

    def copy_from(self, weights_file):
        with open(weights_file):
            pass
        check(lib.caffe_net_copy_trained_from(self._h, weights_file.encode()))

    def save(self, filename):
        check(lib.caffe_net_save(self._h, filename.encode()))

    # ---- B200 extensions
    def skip_outputs(self, names):
        """Declare net outputs this caller never reads (the demo reads `prob` and `loc_pred` only, estimate_pose.py:231): the
        fused plan leaves their heads out of the merged head GEMMs (next_pred is 364 of the 406 head channels) and does not write
        those blobs; forward() stops returning them and reading one raises.  `[]` restores the full net."""
        names = [names] if isinstance(names, str) else list(names)
        check(lib.caffe_net_set_skipped_outputs(self._h, ",".join(names).encode()))
        self._skipped = set(names)

    def set_fusion(self, on):
        check(lib.caffe_net_set_fusion(self._h, int(bool(on))))

    def materialize_intermediates(self, on):
        check(lib.caffe_net_materialize_intermediates(self._h, int(bool(on))))

    @property
    def fused_last_forward(self):
        return bool(lib.caffe_net_fused_last_forward(self._h))

    @property
    def fusion_diagnostic(self):
        return lib.caffe_net_fusion_diagnostic(self._h).decode()

    @property
    def last_forward_launches(self):
        return lib.caffe_net_last_forward_launches(self._h)

    def set_step_timing(self, on):
        check(lib.caffe_net_set_step_timing(self._h, int(bool(on))))

    def step_info(self):
        """[(type, name, ms, flops, bytes)] for the fused steps of the last forward."""
        n = lib.caffe_net_num_steps(self._h)
        if n == 0:
            return []
        names = C.create_string_buffer(256 * n)
        ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
        check(lib.caffe_net_step_info(self._h, names, len(names), ms, fl, by, n))
        rows = names.value.decode().strip().split("\n")
        return [(r.split(" ", 1)[0], r.split(" ", 1)[1], ms[i], fl[i], by[i]) for i, r in enumerate(rows)]

    @property
    def arena_bytes(self):
        return lib.caffe_net_arena_bytes(self._h)

    @property
    def weight_bytes(self):
        return lib.caffe_net_weight_bytes(self._h)

    def set_debug_info(self, on):
        """NetParameter.debug_info: forwards run layer by layer and record mean|x| of every top blob (net.cpp:648-735)."""
        check(lib.caffe_net_set_debug_info(self._h, int(bool(on))))

    def debug_info(self):
        """[(layer, top blob, mean|x|)] of the last forward run with set_debug_info(True)."""
        cap = 8 * max(1, len(self._layer_names))
        names = C.create_string_buffer(1 << 20)
        vals = (C.c_double * cap)()
        n = lib.caffe_net_debug_info(self._h, names, len(names), vals, cap)
        if n < 0:
            raise CaffeError(lib.caffe_last_error().decode(errors="replace"))
        rows = names.value.decode().split("\n")[:n]
        return [tuple(r.split(" ")) + (vals[i],) for i, r in enumerate(rows)]

    def describe_plan(self):
        """The fused execution plan for the current input shapes (steps, L2-resident segments, launch groups, arena),
        planned on the host: works without a GPU."""
        buf = C.create_string_buffer(1 << 20)
        check(lib.caffe_net_describe_plan(self._h, buf, len(buf)))
        return buf.value.decode()

    def set_params(self, weights):
        """weights: {layer name: [arrays]} written through the param views (harness helper)."""
        params = self.params
        for name, arrs in weights.items():
            if name not in params:
                raise KeyError("net has no parameterised layer '%s'" % name)
            if len(arrs) != len(params[name]):
                raise ValueError("layer %s: %d blobs given, %d expected" % (name, len(arrs), len(params[name])))
            for blob, a in zip(params[name], arrs):
                if tuple(blob.shape) != tuple(a.shape):
                    raise ValueError("layer %s: shape %s given, %s expected" % (name, a.shape, blob.shape))
                blob.data[...] = a

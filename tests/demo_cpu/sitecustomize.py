"""TEST INFRASTRUCTURE (tests/test_reference_demo_cpu.py): lets the reference's demo run here, where there is no GPU.

Installs the product's Py3 / SciPy compat layer exactly as a user would (deepcut-cnn_b200/python/compat), then gives the
`caffe` shim's Net.forward a CPU-mode body: the product has no CPU path (Layer::Forward_cpu is LOG(FATAL)), so in CPU mode
-- which the demo selects with --use_cpu -- the forward is computed by THE REFERENCE'S OWN CPU LAYERS (oracle/_ref) on the
shim net's prototxt, parameters and `data` blob, and the outputs are written back into the shim's blobs.  Everything else the
demo touches is the real product: the C++ host's prototxt / caffemodel loaders, Blob reshape, the NumPy data views.
Active only when DC_TEST_CPU_FORWARD=1.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(_ROOT, "deepcut-cnn_b200", "python", "compat"))
import dc_py2compat  # noqa: E402

dc_py2compat.install()

if os.environ.get("DC_TEST_CPU_FORWARD") == "1":
    sys.path.insert(0, _ROOT)
    sys.path.insert(0, os.path.join(_ROOT, "deepcut-cnn_b200", "python"))
    import numpy as np
    import caffe
    from oracle import ref_caffe

    _gpu_forward = caffe.Net.forward

    def _forward(self, blobs=None, start=None, end=None, **kwargs):
        if caffe._caffe.lib.caffe_get_mode() == 1:
            return _gpu_forward(self, blobs=blobs, start=start, end=end, **kwargs)
        assert start is None and end is None and not kwargs
        ref = getattr(self, "_cpu_ref", None)
        if ref is None:
            ref = ref_caffe.RefCaffeNet(open(self._network_file).read())
            ref.set_params({k: [np.array(b.data) for b in bl] for k, bl in self.params.items()})
            self._cpu_ref = ref
        self.reshape()                                   # Net::Reshape: the output blobs take the new input's geometry
        out = ref.forward({"data": np.array(self.blobs["data"].data)}, want=list(self.outputs))
        for k, v in out.items():
            assert tuple(self.blobs[k].shape) == v.shape, (k, self.blobs[k].shape, v.shape)
            self.blobs[k].data[...] = v
        return {k: self.blobs[k].data for k in self.outputs}

    caffe.Net.forward = _forward

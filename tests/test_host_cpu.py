"""Host logic on CPU: prototxt loader, Net::Init graph (names, splits, shapes, outputs), weight
file round trip, error behaviour.  No compute: the product has no CPU forward path."""
import os

import numpy as np
import pytest

import dcutil
from oracle import caffe_ref, prototxt as opt

caffe = dcutil.caffe_module()
REF_PROTOTXT = "/root/reference/models/deepercut/ResNet-152.prototxt"


def test_generated_prototxt_equals_reference_structure():
    if not os.path.exists(REF_PROTOTXT):
        pytest.skip("reference not mounted on this box")
    ref = opt.parse_file(REF_PROTOTXT)
    gen = opt.parse(dcutil.gen_prototxt.generate())
    assert ref == gen           # name, input, input_dim and all 680 layers, field for field


def test_product_loader_reads_the_reference_prototxt_itself():
    # the C++ prototxt loader + Net::Init on the file the reference ships (models/deepercut/ResNet-152.prototxt, incl. its
    # `stride: 1 #2` trailing comments), not on this repo's generated twin: same layers, blobs and parameter shapes as the oracle's
    # independent parser + graph builder derive from that file
    if not os.path.exists(REF_PROTOTXT):
        pytest.skip("reference not mounted on this box")
    net = caffe.Net(REF_PROTOTXT, caffe.TEST)
    onet = caffe_ref.load_net(REF_PROTOTXT)
    assert [n for n in net._layer_names if "_split" not in n] == [opt.get(l, "name") for l in onet.layers]
    assert len(onet.layers) == 680
    assert net.inputs == ["data"] and net.outputs == ["loc_pred", "next_pred", "prob"]
    for name, shape in onet.blob_shapes.items():
        assert tuple(net.blobs[name].shape) == tuple(shape), name
    for lname, shapes in onet.param_shapes.items():
        assert [tuple(b.shape) for b in net.params[lname]] == [tuple(s) for s in shapes], lname
    # the dilated stages as the file states them (res5a_branch2b: 3x3, pad 2, dilation 2 on a stride-1 res5)
    ref = opt.parse_file(REF_PROTOTXT)
    conv = {opt.get(l, "name"): l for l in ref["layer"]}["res5a_branch2b"]["convolution_param"][0]
    assert (conv.get("dilation"), conv.get("pad"), conv.get("stride", [1])) == ([2], [2], [1])
    assert tuple(net.blobs["res5c"].shape[1:2]) == (2048,)


def test_both_parsers_agree():
    txt = dcutil.gen_prototxt.generate(height=64, width=96)
    a, b = opt.parse(txt), dcutil.ptx.parse(txt)
    assert a == b and len(b["layer"]) == 680
    tricky = "a { s: 'x # y' t: \"q\\\"r\" } v: [1, 2.5e-3, -7] w < k: FOO > # trailing"
    assert opt.parse(tricky) == dcutil.ptx.parse(tricky) == {"a": [{"s": ["x # y"], "t": ['q"r']}], "v": [1, 2.5e-3, -7], "w": [{"k": ["FOO"]}]}


def test_net_init_matches_reference_graph(tmp_path):
    path = dcutil.write_prototxt(tmp_path, height=256, width=256)
    net = caffe.Net(path, caffe.TEST)
    # 680 layers + one Split per multiply-read top (insert_splits.cpp): 50 block inputs... counted by the oracle
    onet = caffe_ref.load_net(path)
    assert net.inputs == ["data"]
    assert net.outputs == ["loc_pred", "next_pred", "prob"]          # std::set order, net.cpp:268-274
    assert [n for n in net._layer_names if "_split" not in n] == [opt.get(l, "name") for l in onet.layers]
    for name, shape in onet.blob_shapes.items():
        assert tuple(net.blobs[name].shape) == tuple(shape), name
    # split naming: <blob>_<last writer layer>_<top idx>_split[_k]  (insert_splits.cpp:129-143)
    assert "pool1_pool1_0_split" in net._layer_names
    assert "res2a_res2a_relu_0_split_1" in net._blob_names
    assert "res3b7_res3b7_relu_0_split_4" in net._blob_names          # 5 readers: res4a x2 + 3 skip heads
    # parameter blob orders and shapes
    assert [tuple(b.shape) for b in net.params["bn_conv1"]] == [(64,), (64,), (1,)]
    assert [tuple(b.shape) for b in net.params["scale_conv1"]] == [(64,), (64,)]
    assert [tuple(b.shape) for b in net.params["res5c_up_next"]] == [(2048, 364, 3, 3), (364,)]
    assert [tuple(b.shape) for b in net.params["res5a_branch2b"]] == [(512, 512, 3, 3)]
    for lname, shapes in onet.param_shapes.items():
        assert [tuple(b.shape) for b in net.params[lname]] == [tuple(s) for s in shapes], lname
    # unloaded net: constant-0 fillers, Scale gamma defaults to 1 (scale_layer.cpp:36-40)
    assert not net.params["conv1"][0].data.any() and np.all(net.params["scale_conv1"][0].data == 1)


def test_reshape_propagates(tmp_path):
    path = dcutil.write_prototxt(tmp_path, height=128, width=128)
    net = caffe.Net(path, caffe.TEST)
    net.blobs["data"].reshape(2, 3, 720, 1280)
    net.reshape()
    assert net.blobs["conv1"].shape == (2, 64, 360, 640)
    assert net.blobs["pool1"].shape == (2, 64, 180, 320)
    assert net.blobs["res3b7"].shape == (2, 512, 90, 160)
    assert net.blobs["res5c"].shape == (2, 2048, 45, 80)
    assert net.blobs["res5c_up_pose"].shape == (2, 14, 91, 161)
    assert net.blobs["prob"].shape == (2, 14, 90, 160) and net.blobs["loc_pred"].shape == (2, 28, 90, 160)


def test_insert_splits_known_answer():
    # TestInsertion of the reference (src/caffe/test/test_split_layer.cpp:283-420): data feeds two
    # inner products and a loss reads both -> one split with two outputs, consumers renamed in order.
    src = """
    name: "TestNetwork"
    input: "data" input_dim: 1 input_dim: 3 input_dim: 8 input_dim: 8
    layer { name: "c1" type: "Convolution" bottom: "data" top: "a" convolution_param { num_output: 4 kernel_size: 1 } }
    layer { name: "c2" type: "Convolution" bottom: "data" top: "b" convolution_param { num_output: 4 kernel_size: 1 } }
    layer { name: "r" type: "ReLU" bottom: "a" top: "a" }
    layer { name: "s1" type: "Eltwise" bottom: "a" bottom: "b" top: "s1" }
    layer { name: "s2" type: "Eltwise" bottom: "a" bottom: "s1" top: "s2" }
    """
    out = opt.parse(caffe.insert_splits_text(src))
    names = [opt.get(l, "name") for l in out["layer"]]
    assert names == ["data_input_0_split", "c1", "c2", "r", "a_r_0_split", "s1", "s2"]
    L = {opt.get(l, "name"): l for l in out["layer"]}
    assert L["data_input_0_split"]["top"] == ["data_input_0_split_0", "data_input_0_split_1"]
    assert L["c1"]["bottom"] == ["data_input_0_split_0"] and L["c2"]["bottom"] == ["data_input_0_split_1"]
    assert L["r"]["bottom"] == ["a"] and L["r"]["top"] == ["a"]      # in-place stays in-place
    assert L["a_r_0_split"]["bottom"] == ["a"] and L["a_r_0_split"]["top"] == ["a_r_0_split_0", "a_r_0_split_1"]
    assert L["s1"]["bottom"] == ["a_r_0_split_0", "b"] and L["s2"]["bottom"] == ["a_r_0_split_1", "s1"]


def test_caffemodel_round_trip(tmp_path):
    # python/caffe/test/test_net.py:62-81: save -> load gives identical parameters
    path = dcutil.write_prototxt(tmp_path, stages=(1, 1, 1, 1), height=64, width=64)
    net = caffe.Net(path, caffe.TEST)
    rng = np.random.default_rng(0)
    for name, blobs in net.params.items():
        for b in blobs:
            b.data[...] = rng.standard_normal(b.shape).astype(np.float32)
    model = os.path.join(str(tmp_path), "w.caffemodel")
    net.save(model)
    net2 = caffe.Net(path, model, caffe.TEST)
    for name in net.params:
        for a, b in zip(net.params[name], net2.params[name]):
            assert np.array_equal(a.data, b.data), name
    # wire format sanity: NetParameter.layer is field 100 (tag bytes 0xa2 0x06), BlobProto.data packed floats
    raw = open(model, "rb").read()
    assert b"\xa2\x06" in raw[:64] and len(raw) > 4 * sum(b.count for bl in net.params.values() for b in bl)


def _caffe_subset_messages(packed=True):
    """NetParameter / LayerParameter / BlobProto / BlobShape with the reference's field numbers (src/caffe/proto/caffe.proto:6-22,
    64-100, 310-330), built at run time for the installed google.protobuf: the bytes below are written by the real protobuf
    library, not by this repo's own wire-format writer."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    pkg = "caffe_subset_%s" % ("packed" if packed else "unpacked")
    fd = descriptor_pb2.FileDescriptorProto(name=pkg + ".proto", package=pkg, syntax="proto2")

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, num, typ, label=F.LABEL_OPTIONAL, type_name=None, pack=False):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, num, typ, label
        if type_name:
            f.type_name = "." + pkg + "." + type_name
        if label == F.LABEL_REPEATED and typ in (F.TYPE_FLOAT, F.TYPE_DOUBLE, F.TYPE_INT64):
            f.options.packed = pack
    bs = msg("BlobShape")
    field(bs, "dim", 1, F.TYPE_INT64, F.LABEL_REPEATED, pack=packed)
    bp = msg("BlobProto")
    field(bp, "shape", 7, F.TYPE_MESSAGE, type_name="BlobShape")
    field(bp, "data", 5, F.TYPE_FLOAT, F.LABEL_REPEATED, pack=packed)
    field(bp, "double_data", 8, F.TYPE_DOUBLE, F.LABEL_REPEATED, pack=packed)
    for n, i in (("num", 1), ("channels", 2), ("height", 3), ("width", 4)):
        field(bp, n, i, F.TYPE_INT32)
    lp = msg("LayerParameter")
    field(lp, "name", 1, F.TYPE_STRING)
    field(lp, "type", 2, F.TYPE_STRING)
    field(lp, "bottom", 3, F.TYPE_STRING, F.LABEL_REPEATED)
    field(lp, "top", 4, F.TYPE_STRING, F.LABEL_REPEATED)
    field(lp, "blobs", 7, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name="BlobProto")
    field(lp, "from_the_future", 9999, F.TYPE_STRING)         # a field this reader has never heard of: must be skipped
    field(lp, "future_number", 9998, F.TYPE_FIXED64)
    npm = msg("NetParameter")
    field(npm, "name", 1, F.TYPE_STRING)
    field(npm, "layer", 100, F.TYPE_MESSAGE, F.LABEL_REPEATED, type_name="LayerParameter")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName(pkg + ".NetParameter"))


@pytest.mark.parametrize("packed", [True, False])
def test_caffemodel_written_by_real_protobuf(tmp_path, packed):
    """Net::CopyTrainedLayersFrom (net.cpp:805-858) on a .caffemodel serialised by google.protobuf itself, in every blob encoding the
    reference's Blob::FromProto accepts (blob.cpp:420-470): `shape` + packed float `data`; the deprecated 4-D num/channels/height/
    width header; `double_data`; unpacked repeated scalars (a proto2 parser must take both); unknown fields are skipped; a layer
    the net does not have is ignored."""
    pytest.importorskip("google.protobuf")
    NetMsg = _caffe_subset_messages(packed)
    path = dcutil.write_prototxt(tmp_path, stages=(1, 1, 1, 1), height=64, width=64)
    net = caffe.Net(path, caffe.TEST)
    rng = np.random.default_rng(5)
    want = {}
    m = NetMsg(name="written by protobuf")
    stranger = m.layer.add(name="not_in_this_net", type="Convolution")
    stranger.blobs.add().data.extend([1.0, 2.0])
    for li, (name, blobs) in enumerate(net.params.items()):
        layer = m.layer.add(name=name, type="whatever")
        layer.from_the_future = "x" * 300
        layer.future_number = 7
        want[name] = []
        for b in blobs:
            a = rng.standard_normal(b.shape).astype(np.float32)
            want[name].append(a)
            pb = layer.blobs.add()
            style = li % 3
            if style == 1 and a.ndim <= 4:            # deprecated 4-D header, missing leading axes are 1
                dims = [1] * (4 - a.ndim) + list(a.shape)
                pb.num, pb.channels, pb.height, pb.width = dims
            else:
                pb.shape.dim.extend(a.shape)
            if style == 2:
                pb.double_data.extend(a.astype(np.float64).ravel().tolist())
            else:
                pb.data.extend(a.ravel().tolist())
    model = os.path.join(str(tmp_path), "protobuf_%d.caffemodel" % packed)
    open(model, "wb").write(m.SerializeToString())
    net.copy_from(model)
    for name, arrays in want.items():
        for a, b in zip(arrays, net.params[name]):
            assert np.array_equal(a, b.data), name
    # and the files this repo writes parse back with the real library, bit for bit
    ours = os.path.join(str(tmp_path), "ours.caffemodel")
    net.save(ours)
    back = NetMsg()
    back.ParseFromString(open(ours, "rb").read())
    got = {l.name: l for l in back.layer}
    for name, arrays in want.items():
        assert name in got and len(got[name].blobs) == len(arrays)
        for a, pb in zip(arrays, got[name].blobs):
            assert list(pb.shape.dim) == list(a.shape)
            assert np.array_equal(np.array(pb.data, np.float32).reshape(a.shape), a)


def test_blob_memory_outlives_net(tmp_path):
    # python/caffe/test/test_net.py:48-60
    path = dcutil.write_prototxt(tmp_path, stages=(1, 1, 1, 1), height=64, width=64)
    net = caffe.Net(path, caffe.TEST)
    params = sum(([b for b in bl] for bl in net.params.values()), [])
    blobs = list(net.blobs.values())
    views = [p.data for p in params[:4]] + [b.data for b in blobs[:4]]
    for v in views:
        v[...] = 3.0
    del net, params, blobs
    assert all(float(v.sum()) == 3.0 * v.size for v in views)


def test_host_library_exports_every_symbol_its_header_declares():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "caffe_b200_c.h")).read()
    declared = sorted(set(re.findall(r"\b(caffe_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) > 40
    for name in declared:
        assert hasattr(caffe._caffe.lib, name), "symbol %s declared in caffe_b200_c.h but not exported" % name
    assert set(declared) == set(caffe._caffe.EXPORTS), sorted(set(declared) ^ set(caffe._caffe.EXPORTS))


def test_syncedmem_head_host_side(tmp_path):
    # SyncedMemoryTest.TestCPUWrite / TestInitialization (src/caffe/test/test_syncedmem.cpp:16-50), the transitions that need no
    # device: a fresh blob's memory is UNINITIALIZED, the first host access allocates zeroed memory and moves the head to the CPU
    path = dcutil.write_prototxt(tmp_path, stages=(1, 1, 1, 1), height=64, width=64)
    net = caffe.Net(path, caffe.TEST)
    b = net.blobs["prob"]
    assert b.data_head == "UNINITIALIZED"
    v = b.data
    assert b.data_head == "HEAD_AT_CPU" and not v.any()              # lazily allocated, zero-initialised (syncedmem.cpp:25-33)
    v[...] = 1.0
    assert b.data_head == "HEAD_AT_CPU" and float(b.data.sum()) == v.size
    b.reshape(1, 14, 4, 4)                                             # Reshape within capacity keeps the memory (blob.cpp:23-43)
    assert b.data_head == "HEAD_AT_CPU" and float(b.data.sum()) == 14 * 16


def test_errors_are_exceptions_not_aborts(tmp_path):
    with pytest.raises((IOError, OSError)):
        caffe.Net("/nonexistent/net.prototxt", caffe.TEST)
    bad = os.path.join(str(tmp_path), "bad.prototxt")
    open(bad, "w").write('input: "data" input_dim: 1 input_dim: 3 input_dim: 8 input_dim: 8\n'
                         'layer { name: "x" type: "NoSuchLayer" bottom: "data" top: "y" }')
    with pytest.raises(caffe.CaffeError, match="Unknown layer type"):
        caffe.Net(bad, caffe.TEST)
    open(bad, "w").write('input: "data" input_dim: 1 input_dim: 3 input_dim: 8 input_dim: 8\n'
                         'layer { name: "x" type: "ReLU" bottom: "nope" top: "y" }')
    with pytest.raises(caffe.CaffeError, match="Unknown bottom blob"):
        caffe.Net(bad, caffe.TEST)
    path = dcutil.write_prototxt(tmp_path, stages=(1, 1, 1, 1), height=64, width=64)
    caffe.set_mode_cpu()
    net = caffe.Net(path, caffe.TEST)
    with pytest.raises(caffe.CaffeError, match="no CPU forward path"):
        net.forward()
    with pytest.raises(Exception):
        net.params["conv1"][0].reshape(1, 2, 3)
        net.copy_from("/nonexistent.caffemodel")


def test_shim_covers_every_caffe_name_the_reference_demo_uses():
    """python/pose/estimate_pose.py and pose_demo.py only touch a small pycaffe surface; every attribute they
    use on the `caffe` module, on the Net and on a Blob must exist in the shim (checked by AST, here where the
    reference is mounted; the Python-2 files themselves are not executed)."""
    import ast
    ref = "/root/reference/python/pose"
    if not os.path.isdir(ref):
        pytest.skip("reference not mounted on this box")
    used_module, used_obj = set(), set()
    for fn in ("estimate_pose.py", "pose_demo.py"):
        tree = ast.parse(open(os.path.join(ref, fn)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in ("caffe", "_caffe"):
                used_module.add(node.attr)
            if isinstance(node, ast.Attribute) and node.attr in ("blobs", "forward", "data", "reshape", "params", "inputs", "outputs"):
                used_obj.add(node.attr)
    assert used_module == {"Net", "TEST", "set_mode_gpu", "set_mode_cpu", "set_device"} or used_module <= {"Net", "TEST", "set_mode_gpu", "set_mode_cpu", "set_device"}
    for name in used_module:
        assert hasattr(caffe, name), name
    assert {"blobs", "forward", "data", "reshape"} <= used_obj
    path = dcutil.write_prototxt(__import__("tempfile").mkdtemp(), stages=(1, 1, 1, 1), height=64, width=64)
    net = caffe.Net(path, caffe.TEST)
    blob = net.blobs["data"]
    blob.reshape(1, 3, 40, 48)                                    # estimate_pose.py:227
    blob.data[0, ...] = np.ones((3, 40, 48), np.float32)          # :228
    assert blob.data.shape == (1, 3, 40, 48) and float(blob.data.sum()) == 3 * 40 * 48
    assert callable(net.forward) and "loc_pred" in net.blobs and "prob" in net.blobs      # :229-232


def test_prototxt_rejects_unknown_fields_like_protobuf_textformat():
    """ReadProtoFromTextFile (io.cpp:34-44) is protobuf TextFormat: a field the schema lacks is an error, not a note -- a
    misspelt parameter must not silently build a different net.  Only caffe.proto fields of layer types outside the forward
    path (and the training-only propagate_down) are skipped."""
    caffe = dcutil.caffe_module()
    caffe.set_mode_cpu()
    head = 'name: "t" input: "data" input_dim: 1 input_dim: 3 input_dim: 8 input_dim: 8 '
    bad = head + 'layer { name: "c" type: "Convolution" bottom: "data" top: "c" convolution_param { num_output: 4 kernel_size: 1 strde: 2 } }'
    with pytest.raises(caffe._caffe.CaffeError, match="ConvolutionParameter has no field named 'strde'"):
        caffe.Net.from_string(bad, caffe.TEST)
    with pytest.raises(caffe._caffe.CaffeError, match="no field named 'inptu'"):
        caffe.Net.from_string(head.replace('input: "data"', 'inptu: "data"'), caffe.TEST)
    ok = head + ('layer { name: "c" type: "Convolution" bottom: "data" top: "c" propagate_down: false dropout_param { dropout_ratio: 0.5 } '
                 'convolution_param { num_output: 4 kernel_size: 1 stride: 2 } }')
    net = caffe.Net.from_string(ok, caffe.TEST)
    assert net.blobs["c"].shape == (1, 4, 4, 4)

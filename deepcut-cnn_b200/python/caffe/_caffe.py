"""ctypes binding of include/caffe_b200_c.h (the role of python/caffe/_caffe.cpp in the reference)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(os.path.dirname(_HERE))
_LIB_PATH = os.path.join(_PKG, "libcaffe_b200.so")

TRAIN, TEST = 0, 1


class CaffeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError("%s is missing: run __graft_entry__.build()" % _LIB_PATH)
    # libcaffe_b200 resolves libdeepcut_b200 through its $ORIGIN rpath
    lib = C.CDLL(_LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, ci, cs = C.c_void_p, C.c_int, C.c_char_p
    sig = {
        "caffe_last_error": (cs, []), "caffe_set_mode": (ci, [ci]), "caffe_get_mode": (ci, []),
        "caffe_set_device": (ci, [ci]), "caffe_device_count": (ci, []), "caffe_set_log_level": (None, [ci]),
        "caffe_stream": (vp, []), "caffe_sync": (ci, []),
        "caffe_net_create": (vp, [cs, ci]), "caffe_net_create_from_string": (vp, [cs, ci]),
        "caffe_net_destroy": (None, [vp]), "caffe_net_copy_trained_from": (ci, [vp, cs]),
        "caffe_net_save": (ci, [vp, cs]), "caffe_net_forward": (ci, [vp]),
        "caffe_net_forward_from_to": (ci, [vp, ci, ci]), "caffe_net_reshape": (ci, [vp]),
        "caffe_net_name": (cs, [vp]), "caffe_net_num_blobs": (ci, [vp]), "caffe_net_blob_name": (cs, [vp, ci]),
        "caffe_net_num_layers": (ci, [vp]), "caffe_net_layer_name": (cs, [vp, ci]),
        "caffe_net_layer_type": (cs, [vp, ci]), "caffe_net_layer_num_blobs": (ci, [vp, ci]),
        "caffe_net_layer_num_bottoms": (ci, [vp, ci]), "caffe_net_layer_bottom_id": (ci, [vp, ci, ci]),
        "caffe_net_layer_num_tops": (ci, [vp, ci]), "caffe_net_layer_top_id": (ci, [vp, ci, ci]),
        "caffe_net_num_inputs": (ci, [vp]), "caffe_net_input_index": (ci, [vp, ci]),
        "caffe_net_num_outputs": (ci, [vp]), "caffe_net_output_index": (ci, [vp, ci]),
        "caffe_net_layer_weights_changed": (ci, [vp, ci]),
        "caffe_net_blob": (vp, [vp, ci]), "caffe_net_layer_blob": (vp, [vp, ci, ci]),
        "caffe_blob_release": (None, [vp]), "caffe_blob_num_axes": (ci, [vp]), "caffe_blob_shape": (ci, [vp, ci]),
        "caffe_blob_count": (ci, [vp]), "caffe_blob_reshape": (ci, [vp, ci, C.POINTER(ci)]),
        "caffe_blob_mutable_cpu_data": (vp, [vp]), "caffe_blob_cpu_data": (vp, [vp]),
        "caffe_blob_mutable_cpu_diff": (vp, [vp]), "caffe_blob_gpu_data": (vp, [vp]),
        "caffe_blob_mutable_gpu_data": (vp, [vp]), "caffe_blob_overwrite_gpu_data": (vp, [vp]), "caffe_blob_data_head": (ci, [vp]),
        "caffe_net_set_fusion": (ci, [vp, ci]), "caffe_net_materialize_intermediates": (ci, [vp, ci]),
        "caffe_net_fused_last_forward": (ci, [vp]), "caffe_net_fusion_diagnostic": (cs, [vp]),
        "caffe_net_last_forward_launches": (C.c_longlong, [vp]),
        "caffe_insert_splits_text": (ci, [cs, C.c_char_p, ci]),
        "caffe_net_set_step_timing": (ci, [vp, ci]), "caffe_net_num_steps": (ci, [vp]),
        "caffe_net_step_info": (ci, [vp, C.c_char_p, ci, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), ci]),
        "caffe_net_arena_bytes": (C.c_longlong, [vp]), "caffe_net_weight_bytes": (C.c_longlong, [vp]),
        "caffe_net_describe_plan": (ci, [vp, C.c_char_p, ci]), "caffe_net_blob_fresh": (ci, [vp, ci]),
        "caffe_net_set_skipped_outputs": (ci, [vp, cs]),
        "caffe_net_set_debug_info": (ci, [vp, ci]), "caffe_net_debug_info": (ci, [vp, C.c_char_p, ci, C.POINTER(C.c_double), ci]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTS = _load()


def check(rc):
    if rc != 0:
        raise CaffeError(lib.caffe_last_error().decode(errors="replace"))


def check_ptr(p):
    if not p:
        raise CaffeError(lib.caffe_last_error().decode(errors="replace"))
    return p


def set_mode_cpu():
    check(lib.caffe_set_mode(0))


def set_mode_gpu():
    check(lib.caffe_set_mode(1))


def set_device(device_id):
    check(lib.caffe_set_device(int(device_id)))


def device_count():
    return lib.caffe_device_count()


def set_log_level(level):
    lib.caffe_set_log_level(int(level))


def sync():
    check(lib.caffe_sync())


def insert_splits_text(prototxt_text):
    buf = C.create_string_buffer(1 << 22)
    check(lib.caffe_insert_splits_text(prototxt_text.encode(), buf, len(buf)))
    return buf.value.decode()

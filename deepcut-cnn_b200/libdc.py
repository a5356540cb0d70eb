"""ctypes binding of include/deepcut_b200.h.  Fails loudly if the CUDA library is missing:
there is no CPU fallback for the hot path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdeepcut_b200.so")


class ConvArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("cin", C.c_int),
                ("cout", C.c_int), ("kh", C.c_int), ("kw", C.c_int), ("pad", C.c_int), ("dilation", C.c_int),
                ("w_packed", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("residual", C.c_void_p),
                ("relu", C.c_int), ("out_f32_rows", C.c_int), ("ldc", C.c_int), ("out", C.c_void_p), ("stride", C.c_int),
                ("splitk_workspace", C.c_void_p), ("splitk_workspace_bytes", C.c_size_t),
                ("x_plane", C.c_longlong), ("out_plane", C.c_longlong), ("residual_plane", C.c_longlong),
                ("weights_evict_last", C.c_int)]


_SIGS = {
    "dc_version": (C.c_int, []),
    "dc_last_error": (C.c_char_p, []),
    "dc_device_count": (C.c_int, []),
    "dc_init": (C.c_int, [C.c_int]),
    "dc_launch_count": (C.c_longlong, []),
    "dc_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "dc_free": (C.c_int, [C.c_void_p]),
    "dc_malloc_host": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "dc_free_host": (C.c_int, [C.c_void_p]),
    "dc_memcpy_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "dc_memset_async": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]),
    "dc_stream_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "dc_stream_destroy": (C.c_int, [C.c_void_p]),
    "dc_stream_sync": (C.c_int, [C.c_void_p]),
    "dc_device_sync": (C.c_int, []),
    "dc_mem_info": (C.c_int, [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "dc_event_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "dc_event_destroy": (C.c_int, [C.c_void_p]),
    "dc_event_record": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dc_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "dc_event_sync": (C.c_int, [C.c_void_p]),
    "dc_images_u8_to_blob": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p, C.c_void_p]),
    "dc_graph_begin": (C.c_int, [C.c_void_p]),
    "dc_graph_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "dc_graph_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dc_graph_destroy": (C.c_int, [C.c_void_p]),
    "dc_fold_bn_scale": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p]),
    "dc_packed_rows": (C.c_int, [C.c_int]),
    "dc_tile_n": (C.c_int, [C.c_int]),
    "dc_pack_conv_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dc_pack_deconv_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dc_pack_conv1_weight": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dc_conv_forward": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "dc_set_split_k": (C.c_int, [C.c_int]),
    "dc_get_split_k": (C.c_int, []),
    "dc_set_reserved_sms": (C.c_int, [C.c_int]),
    "dc_get_reserved_sms": (C.c_int, []),
    "dc_splitk_workspace_bytes": (C.c_size_t, []),
    "dc_set_split_k_min_steps": (C.c_int, [C.c_int]),
    "dc_get_split_k_min_steps": (C.c_int, []),
    "dc_conv1_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "dc_conv1_tc_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dc_pack_conv1_tc_weight": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dc_conv1_tc_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "dc_maxpool_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "dc_pool_out_size": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "dc_subsample_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p]),
    "dc_head_finish": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dc_pose_from_maps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p]),
    "dc_preprocess_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_void_p)]),
    "dc_preprocess_plan_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "dc_preprocess_u8_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p]),
    "dc_preprocess_plan_destroy": (C.c_int, [C.c_void_p]),
    "dc_nchw_to_split": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dc_split_to_nchw": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dc_bn_forward_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dc_scale_forward_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dc_relu_forward": (C.c_int, [C.c_void_p, C.c_longlong, C.c_float, C.c_void_p, C.c_void_p]),
    "dc_sigmoid_forward": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "dc_axpby_forward": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_longlong, C.c_void_p, C.c_void_p]),
    "dc_crop_forward_nchw": (C.c_int, [C.c_void_p] + [C.c_int] * 8 + [C.c_void_p, C.c_void_p]),
    "dc_maxpool_forward_nchw": (C.c_int, [C.c_void_p] + [C.c_int] * 12 + [C.c_void_p, C.c_void_p]),
    "dc_conv_direct_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 13 + [C.c_void_p, C.c_void_p]),
    "dc_deconv_direct_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 13 + [C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run __graft_entry__.build() (the hot path has no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)      # AttributeError = header/library mismatch
            fn.restype = res
            fn.argtypes = args
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def check(rc):
    if rc != 0:
        raise RuntimeError("libdeepcut_b200: rc=%d: %s" % (rc, lib().dc_last_error().decode()))

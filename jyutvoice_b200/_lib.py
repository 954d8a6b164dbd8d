"""ctypes binding of include/jyutvoice_b200.h.  There is no CPU fallback: if the CUDA library is
missing or fails to load, importing the product path raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjyutvoice_b200.so")

JV_OK = 0
JV_ERR_INVALID = -1
JV_ERR_CUDA = -2
JV_ERR_STATE = -3
PREC = {"fp32": 0, "bf16": 1}

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t
c_int64 = ctypes.c_int64
P_i32 = ctypes.POINTER(ctypes.c_int32)
P_i64 = ctypes.POINTER(ctypes.c_int64)
P_f32 = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes); this table is also what tests check against the header
SIGNATURES = {
    "jv_version": (c_int, []),
    "jv_last_error": (ctypes.c_char_p, []),
    "jv_launch_count": (ctypes.c_uint64, []),
    "jv_graph_launch_count": (ctypes.c_uint64, []),
    "jv_simt_fallback_count": (ctypes.c_uint64, []),
    "jv_estimator_set_stream_format": (c_int, [c_void_p, c_int]),
    "jv_estimator_saturation_count": (c_int, [c_void_p, c_int, P_i64]),
    "jv_estimator_time_embedding": (c_int, [c_void_p, P_f32, c_int, c_void_p, c_void_p]),
    "jv_hift_stft": (c_int, [c_void_p, c_int, c_int, P_i32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "jv_estimator_create": (c_int, [c_int, c_int, ctypes.POINTER(c_void_p)]),
    "jv_estimator_destroy": (None, [c_void_p]),
    "jv_estimator_set_weight": (c_int, [c_void_p, ctypes.c_char_p, c_void_p, P_i64, c_int]),
    "jv_estimator_finalize": (c_int, [c_void_p]),
    "jv_estimator_set_chunk": (c_int, [c_void_p, c_int]),
    "jv_cfm_workspace_bytes": (c_size_t, [c_void_p, c_int, P_i32]),
    "jv_cfm_solve_workspace_bytes": (c_size_t, [c_void_p, c_int, P_i32]),
    "jv_estimator_forward": (c_int, [c_void_p, c_int, c_int, P_i32, c_void_p, c_void_p, P_f32, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    "jv_cfm_solve": (c_int, [c_void_p, c_int, c_int, P_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float,
                             c_int, P_f32, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "jv_hift_create": (c_int, [c_int, c_int, ctypes.POINTER(c_void_p)]),
    "jv_hift_destroy": (None, [c_void_p]),
    "jv_hift_set_weight": (c_int, [c_void_p, ctypes.c_char_p, c_void_p, P_i64, c_int]),
    "jv_hift_finalize": (c_int, [c_void_p]),
    "jv_hift_workspace_bytes": (c_size_t, [c_void_p, c_int, P_i32]),
    "jv_hift_f0": (c_int, [c_void_p, c_int, c_int, P_i32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "jv_hift_source": (c_int, [c_void_p, c_int, c_int, P_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "jv_hift_decode": (c_int, [c_void_p, c_int, c_int, P_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "jv_text_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "jv_text_destroy": (None, [c_void_p]),
    "jv_text_set_weight": (c_int, [c_void_p, ctypes.c_char_p, c_void_p, P_i64, c_int]),
    "jv_text_finalize": (c_int, [c_void_p]),
    "jv_text_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, P_i32]),
    "jv_text_encode": (c_int, [c_void_p, c_int, c_int, P_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "jv_text_durations": (c_int, [c_void_p, c_int, c_int, P_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "jv_length_durations": (c_int, [c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p]),
    "jv_length_align": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "jv_flowenc_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "jv_flowenc_destroy": (None, [c_void_p]),
    "jv_flowenc_set_weight": (c_int, [c_void_p, ctypes.c_char_p, c_void_p, P_i64, c_int]),
    "jv_flowenc_finalize": (c_int, [c_void_p]),
    "jv_flowenc_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, P_i32]),
    "jv_flowenc_encode": (c_int, [c_void_p, c_int, c_int, P_i32, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    "jv_profile_begin": (c_int, []),
    "jv_profile_end": (c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    "jv_bench_gemm": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_double)]),
    "jv_test_gemm": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m jyutvoice_b200.build` "
                "(jyutvoice_b200 has no CPU or PyTorch fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError here = the .so is older than the header: rebuild
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc == JV_OK:
        return
    msg = lib().jv_last_error().decode("utf-8", "replace")
    if rc == JV_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)


def i32_array(values):
    arr = (ctypes.c_int32 * len(values))(*[int(v) for v in values])
    return arr


def f32_array(values):
    return (ctypes.c_float * len(values))(*[float(v) for v in values])


def set_weights(handle, setter, items):
    """items: iterable of (key, tensor fp32 contiguous, cpu or cuda)."""
    for key, t in items:
        t = t.detach().contiguous().float()
        shape = (ctypes.c_int64 * t.dim())(*t.shape)
        check(setter(handle, key.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim()))

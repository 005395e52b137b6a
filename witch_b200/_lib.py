"""ctypes binding of libwitch_b200.so (include/witch_b200.h). Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libwitch_b200.so")

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u64p = ctypes.POINTER(ctypes.c_uint64)

SYMBOLS = {
    "witch_last_error": (ctypes.c_char_p, []),
    "witch_version": (ctypes.c_char_p, []),
    "witch_device_count": (ctypes.c_int, []),
    "witch_ehmm_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_void_p)]),
    "witch_ehmm_create_cached": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.c_char_p, ctypes.POINTER(ctypes.c_int),
                                                ctypes.POINTER(ctypes.c_void_p)]),
    "witch_ehmm_destroy": (None, [ctypes.c_void_p]),
    "witch_ehmm_count": (ctypes.c_int, [ctypes.c_void_p]),
    "witch_ehmm_alphabet": (ctypes.c_int, [ctypes.c_void_p]),
    "witch_ehmm_info": (ctypes.c_int, [ctypes.c_void_p, c_i32p, c_i32p]),
    "witch_queries_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, c_i64p, ctypes.POINTER(ctypes.c_void_p)]),
    "witch_queries_destroy": (None, [ctypes.c_void_p]),
    "witch_queries_count": (ctypes.c_int, [ctypes.c_void_p]),
    "witch_score": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, c_f32p, c_u8p, c_f32p, c_u8p]),
    "witch_score_dev": (ctypes.c_int, [ctypes.c_void_p] * 7),
    "witch_weights_topk": (ctypes.c_int, [ctypes.c_void_p, c_f32p, c_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i32p, c_f64p, c_i32p]),
    "witch_weights_topk_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "witch_align": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_i32p, c_i32p, c_i64p, c_i32p]),
    "witch_align_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_i32p, c_i32p, c_i64p, ctypes.c_void_p, ctypes.c_void_p]),
    "witch_graph_align": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i32p, c_i64p, ctypes.c_char_p, c_i32p, c_i32p, c_f64p, c_i64p,
                                         c_i32p, ctypes.c_int, c_i64p, c_i32p, c_i32p, ctypes.c_int, c_i64p, ctypes.c_void_p, c_i32p]),
    "witch_merge_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_i64p, c_i32p, c_u8p, ctypes.c_char_p, ctypes.c_int, c_i32p,
                                        c_i64p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "witch_kernel_launches": (ctypes.c_uint64, []),
    "witch_prof_enable": (None, [ctypes.c_int]),
    "witch_prof_reset": (None, []),
    "witch_prof_get": (ctypes.c_double, [ctypes.c_int, c_f64p, c_u64p]),
    "witch_measure_fp32_peak": (ctypes.c_double, [ctypes.c_double]),
    "witch_debug_fwdbwd": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_i32p, c_i32p, ctypes.c_int, c_f32p, c_f32p]),
}

_lib = None


class WitchError(RuntimeError):
    pass


def load():
    """Load the shared library and declare every symbol of include/witch_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WitchError("libwitch_b200.so is not built (%s); run `python -m witch_b200.build` "
                         "-- there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise WitchError("witch_b200 error %d: %s" % (rc, load().witch_last_error().decode()))

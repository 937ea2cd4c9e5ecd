"""ctypes binding of include/ringo_cuda.h (libringo_b200.so).  Fails loudly when the
library is missing -- there is deliberately no fallback path."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RINGO_B200_LIB", os.path.join(HERE, "libringo_b200.so"))   # env: tuning variants only

RS_OK = 0
RS_TERM_MAX_W = 8

c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_size = ctypes.c_size_t
c_vp = ctypes.c_void_p


class RingoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ringo_b200 error {code}: {msg}")
        self.code = code


class Term(ctypes.Structure):
    _fields_ = [("constraint", ctypes.c_int32), ("coeff", ctypes.c_int32), ("pw", ctypes.c_int32),
                ("n_w", ctypes.c_int32), ("w", ctypes.c_int32 * RS_TERM_MAX_W)]


# name -> (restype, argtypes); mirrors include/ringo_cuda.h one to one
_SIGS = {
    "rs_init": (ctypes.c_int, [ctypes.c_int]),
    "rs_set_stream": (ctypes.c_int, [c_vp]),
    "rs_sync": (ctypes.c_int, []),
    "rs_last_error": (ctypes.c_char_p, []),
    "rs_version": (ctypes.c_char_p, []),
    "rs_launch_count": (ctypes.c_ulonglong, []),
    "rs_bench_imad_peak": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "rs_pinned_alloc": (c_vp, [c_size]),
    "rs_pinned_free": (None, [c_vp]),
    "rs_field_create": (ctypes.c_int, [c_u64p, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "rs_field_destroy": (None, [c_vp]),
    "rs_field_limbs": (ctypes.c_int, [c_vp]),
    "rs_field_to_mont": (ctypes.c_int, [c_vp, c_u64p, c_u64p]),
    "rs_field_from_mont": (ctypes.c_int, [c_vp, c_u64p, c_u64p]),
    "rs_batch_alloc": (ctypes.c_int, [c_vp, c_size, c_size, ctypes.POINTER(c_vp)]),
    "rs_batch_wrap": (ctypes.c_int, [c_vp, c_vp, c_size, c_size, ctypes.POINTER(c_vp)]),
    "rs_batch_free": (None, [c_vp]),
    "rs_batch_device_ptr": (c_vp, [c_vp]),
    "rs_batch_polys": (c_size, [c_vp]),
    "rs_batch_rank": (c_size, [c_vp]),
    "rs_batch_upload": (ctypes.c_int, [c_vp, c_size, c_size, c_vp]),
    "rs_batch_download": (ctypes.c_int, [c_vp, c_size, c_size, c_vp]),
    "rs_batch_copy": (ctypes.c_int, [c_vp, c_size, c_size, c_vp, c_size, c_size, c_size]),
    "rs_batch_zero": (ctypes.c_int, [c_vp, c_size, c_size, c_size]),
    "rs_ntt_create": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "rs_ntt_destroy": (None, [c_vp]),
    "rs_ntt_rank_log2": (ctypes.c_int, [c_vp]),
    "rs_ntt_tables": (ctypes.c_int, [c_vp, c_u64p, c_u64p, c_u64p]),
    "rs_ntt_forward": (ctypes.c_int, [c_vp, c_vp, c_size, c_vp, c_size, c_size]),
    "rs_ntt_inverse": (ctypes.c_int, [c_vp, c_vp, c_size, c_vp, c_size, c_size]),
    "rs_ntt_profile": (ctypes.c_int, [c_vp, c_vp, c_size, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]),
    "rs_ntt_inverse_embed": (ctypes.c_int, [c_vp, c_vp, c_size, c_vp, c_size, c_size, ctypes.c_int]),
    "rs_vec_add": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_vp, c_size, c_size]),
    "rs_vec_sub": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_vp, c_size, c_size]),
    "rs_vec_neg": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_size]),
    "rs_vec_mul": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_vp, c_size, c_size]),
    "rs_vec_mul_add": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_vp, c_size, c_size]),
    "rs_vec_mul_sub": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_vp, c_size, c_size]),
    "rs_vec_scalar_mul": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_u64p, c_size]),
    "rs_vec_scalar_mul_add": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_u64p, c_size]),
    "rs_vec_scalar_mul_sub": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_u64p, c_size]),
    "rs_eval_circuit": (ctypes.c_int, [c_vp, c_size, c_vp, c_vp, ctypes.POINTER(Term), c_size, c_u64p, c_size, c_u64p]),
    "rs_quorem_vanishing": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_vp, c_size, c_size]),
    "rs_aut": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_size, ctypes.c_long, ctypes.c_int]),
    "rs_reverse_scale": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_u64p]),
    "rs_poly_evaluate": (ctypes.c_int, [c_vp, c_size, c_size, c_u64p, c_u64p]),
}

_lib = None


def load():
    """dlopen libringo_b200.so and bind every symbol the header declares."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python ringo-snark_b200/build.py` "
                          "(there is no CPU fallback for the CUDA path)")
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(L, name)      # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def lib():
    return load()


def declared_symbols():
    return sorted(_SIGS)


def check(rc):
    if rc != RS_OK:
        raise RingoError(rc, lib().rs_last_error().decode())
    return rc

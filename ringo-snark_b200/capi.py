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


RS_JINDO_MAX_LIMBS = 4
RS_JINDO_MAX_EXP = 64


class JindoParams(ctypes.Structure):
    """rs_jindo_params (include/ringo_jindo.h)"""
    _fields_ = [("d", ctypes.c_int32), ("slots", ctypes.c_int32), ("exp", ctypes.c_int32), ("base", ctypes.c_uint64),
                ("rows", ctypes.c_int32), ("cols", ctypes.c_int32),
                ("in_msis", ctypes.c_int32), ("mlwe", ctypes.c_int32), ("out_msis", ctypes.c_int32),
                ("log_in_cut", ctypes.c_int32), ("log_out_cut", ctypes.c_int32),
                ("n_q", ctypes.c_int32), ("n_qout", ctypes.c_int32),
                ("q", ctypes.c_uint64 * RS_JINDO_MAX_LIMBS), ("qout", ctypes.c_uint64 * RS_JINDO_MAX_LIMBS),
                ("psi_q", ctypes.c_uint64 * RS_JINDO_MAX_LIMBS), ("psi_qout", ctypes.c_uint64 * RS_JINDO_MAX_LIMBS),
                ("delta_inv", ctypes.c_double * RS_JINDO_MAX_EXP),
                ("cdt_len", ctypes.c_int32), ("cdt_tail_lo", ctypes.c_int64), ("cdt_sigma", ctypes.c_double)]


class JindoTape(ctypes.Structure):
    """rs_jindo_tape (include/ringo_jindo.h)"""
    _fields_ = [("first_row", c_vp), ("last_row", c_vp), ("mask_rows", c_vp), ("n_mask_rows", c_size),
                ("twin_u", c_vp), ("n_twin_u", c_size), ("cosac_g", c_vp), ("n_cosac_polys", c_size), ("mlwe", c_vp),
                ("on_device", ctypes.c_int32), ("mlwe_mask_only", ctypes.c_int32)]


class JindoProof(ctypes.Structure):
    """rs_jindo_proof (include/ringo_jindo.h)"""
    _fields_ = [("partial", c_vp), ("partial_mask", c_vp), ("encode", c_vp), ("mlwe", c_vp), ("incommit", c_vp)]


class JindoNoise(ctypes.Structure):
    """rs_jindo_noise (include/ringo_sampler.h)"""
    _fields_ = [("ecd_blind_std", ctypes.c_double), ("mask_blind_std", ctypes.c_double), ("mask_std", ctypes.c_double),
                ("mask_mlwe_std", ctypes.c_double)]


class JindoDerived(ctypes.Structure):
    """rs_jindo_derived (include/ringo_jindo.h)"""
    _fields_ = [("p", JindoParams), ("noise", JindoNoise), ("ecd_std", ctypes.c_double), ("mlwe_std", ctypes.c_double),
                ("res_two_nm", ctypes.c_double), ("in_com_dcmp_two_nm", ctypes.c_double), ("com_size", ctypes.c_double),
                ("pf_size", ctypes.c_double), ("rank", ctypes.c_uint64), ("batch", ctypes.c_int32), ("in_com_dcmp_len", ctypes.c_int32),
                ("q_bits", ctypes.c_int32), ("qout_bits", ctypes.c_int32)]


class LinChecker(ctypes.Structure):
    """rs_lin_checker (include/ringo_buckler.h)"""
    _fields_ = [("kind", ctypes.c_int32), ("is_ntt", ctypes.c_int32), ("idx_inv", ctypes.c_int64), ("pairs", c_vp), ("n_pairs", c_size)]


class BucklerCircuit(ctypes.Structure):
    """rs_buckler_circuit (include/ringo_buckler.h)"""
    _fields_ = [("rank", c_size), ("emb_rank", c_size), ("w_cnt", c_size), ("pw_cnt", c_size),
                ("arith_max_rank", c_size), ("sum_max_rank", c_size),
                ("arith_terms", c_vp), ("n_arith_terms", c_size), ("arith_coeffs", c_vp), ("n_arith_coeffs", c_size),
                ("sum_terms", c_vp), ("n_sum_terms", c_size), ("sum_coeffs", c_vp), ("n_sum_coeffs", c_size),
                ("lin", c_vp), ("n_lin", c_size)]


class BucklerRand(ctypes.Structure):
    """rs_buckler_rand (include/ringo_buckler.h)"""
    _fields_ = [("blind", c_vp), ("lin_mask", c_vp), ("lin_mask_sum", c_vp), ("sum_mask", c_vp), ("sum_mask_sum", c_vp),
                ("tapes", c_vp), ("n_tapes", c_size), ("tapes_pinned", ctypes.c_int32)]


class BucklerProof(ctypes.Structure):
    """rs_buckler_proof (include/ringo_buckler.h)"""
    _fields_ = [("witness", c_vp), ("lin_mask_sum", c_vp), ("sum_mask_sum", c_vp), ("evals", c_vp), ("partial", c_vp),
                ("partial_mask", c_vp), ("encode", c_vp), ("mlwe", c_vp), ("incommit", c_vp), ("eval_point", c_vp)]


# name -> (restype, argtypes); mirrors include/ringo_cuda.h, ringo_jindo.h, ringo_buckler.h and ringo_sampler.h one to one
_SIGS = {
    "rs_init": (ctypes.c_int, [ctypes.c_int]),
    "rs_set_stream": (ctypes.c_int, [c_vp]),
    "rs_sync": (ctypes.c_int, []),
    "rs_get_stream": (c_vp, []),
    "rs_stream_create": (ctypes.c_int, [ctypes.POINTER(c_vp)]),
    "rs_stream_destroy": (None, [c_vp]),
    "rs_stream_wait": (ctypes.c_int, [c_vp, c_vp]),
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
    "rs_field_mul": (ctypes.c_int, [c_vp, c_u64p, c_u64p, c_u64p]),
    "rs_field_add": (ctypes.c_int, [c_vp, c_u64p, c_u64p, c_u64p]),
    "rs_field_sub": (ctypes.c_int, [c_vp, c_u64p, c_u64p, c_u64p]),
    "rs_batch_alloc": (ctypes.c_int, [c_vp, c_size, c_size, ctypes.POINTER(c_vp)]),
    "rs_batch_wrap": (ctypes.c_int, [c_vp, c_vp, c_size, c_size, ctypes.POINTER(c_vp)]),
    "rs_batch_free": (None, [c_vp]),
    "rs_batch_device_ptr": (c_vp, [c_vp]),
    "rs_batch_polys": (c_size, [c_vp]),
    "rs_batch_rank": (c_size, [c_vp]),
    "rs_batch_upload": (ctypes.c_int, [c_vp, c_size, c_size, c_vp]),
    "rs_batch_download": (ctypes.c_int, [c_vp, c_size, c_size, c_vp]),
    "rs_batch_download_async": (ctypes.c_int, [c_vp, c_size, c_size, c_vp]),
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
    "rs_vec_powers": (ctypes.c_int, [c_vp, c_size, c_size, c_u64p]),
    "rs_field_set_bytes": (ctypes.c_int, [c_vp, c_vp, c_size, c_u64p]),
    "rs_field_marshal": (ctypes.c_int, [c_vp, c_u64p, c_vp]),
    "rs_opening_download_all": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp]),
    "rs_rand_encode_blind": (ctypes.c_int, [c_vp, c_size, c_size, c_size, c_u64p]),
    "rs_eval_circuit": (ctypes.c_int, [c_vp, c_size, c_vp, c_vp, ctypes.POINTER(Term), c_size, c_u64p, c_size, c_u64p]),
    "rs_quorem_vanishing": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_vp, c_size, c_size]),
    "rs_aut": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_size, ctypes.c_long, ctypes.c_int]),
    "rs_reverse_scale": (ctypes.c_int, [c_vp, c_size, c_vp, c_size, c_u64p]),
    "rs_poly_evaluate": (ctypes.c_int, [c_vp, c_size, c_size, c_u64p, c_u64p]),
    "rs_poly_evaluate_many": (ctypes.c_int, [c_vp, c_vp, c_vp, c_size, c_u64p, c_u64p]),
    # ringo_jindo.h
    "rs_jindo_create": (ctypes.c_int, [c_vp, ctypes.POINTER(JindoParams), c_vp, c_vp, c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "rs_jindo_create_crs": (ctypes.c_int, [c_vp, ctypes.POINTER(JindoParams), c_vp, c_vp, c_vp, ctypes.POINTER(c_vp)]),
    "rs_jindo_new_parameters": (ctypes.c_int, [c_u64p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(JindoDerived)]),
    "rs_jindo_cdt_tables": (ctypes.c_int, [ctypes.POINTER(JindoDerived), c_vp]),
    "rs_jindo_get_params": (ctypes.c_int, [c_vp, ctypes.POINTER(JindoParams)]),
    "rs_jindo_commit_key_download": (ctypes.c_int, [c_vp, ctypes.c_int, c_vp]),
    "rs_jindo_destroy": (None, [c_vp]),
    "rs_jindo_set_twin_stream": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "rs_jindo_twin_stream_draw": (ctypes.c_int, [c_vp, c_size, c_vp]),
    "rs_jindo_set_mlwe_stream": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int64]),
    "rs_jindo_set_cosac_stream": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    "rs_jindo_cosac_status": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_int)]),
    "rs_jindo_cosac_sample": (ctypes.c_int, [c_vp, c_vp, c_vp, c_size, c_vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_int)]),
    "rs_jindo_mlwe_stream_sample": (ctypes.c_int, [c_vp, c_size, c_vp]),
    "rs_jindo_commit": (ctypes.c_int, [c_vp, c_vp, c_size, c_size, ctypes.POINTER(JindoTape), ctypes.POINTER(c_vp), c_vp]),
    "rs_jindo_commit_async": (ctypes.c_int, [c_vp, c_vp, c_size, c_size, ctypes.POINTER(JindoTape), ctypes.POINTER(c_vp), c_vp]),
    "rs_jindo_commit_host": (ctypes.c_int, [c_vp, c_vp, c_size, ctypes.POINTER(JindoTape), ctypes.POINTER(c_vp), c_vp]),
    "rs_jindo_commit_columns": (ctypes.c_int, [c_vp, c_vp, c_size, c_size, ctypes.POINTER(JindoTape), ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "rs_jindo_outer_commit": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "rs_opening_incommit_device": (c_vp, [c_vp]),
    "rs_jindo_shard_columns": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "rs_jindo_outer_commit_sharded": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_vp]),
    "rs_jindo_allgather_incommit": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, ctypes.c_int, c_vp]),
    "rs_nccl_unique_id": (ctypes.c_int, [c_vp]),
    "rs_nccl_comm_create": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "rs_nccl_comm_destroy": (None, [c_vp]),
    "rs_allgather_u64": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_size, c_vp]),
    "rs_opening_free": (None, [c_vp]),
    "rs_opening_download": (ctypes.c_int, [c_vp, ctypes.c_int, c_size, c_size, c_vp]),
    "rs_jindo_partial": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "rs_jindo_respond": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "rs_jindo_partial_at": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "rs_jindo_respond_bytes": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "rs_jindo_left_vec": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "rs_jindo_encode_challenges": (ctypes.c_int, [c_vp, c_vp, c_size, ctypes.c_int, c_vp]),
    "rs_opening_mul_add_bytes": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "rs_opening_combine_bytes": (ctypes.c_int, [c_vp, c_vp, c_size, c_vp, ctypes.POINTER(c_vp)]),
    "rs_jindo_verify": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.c_int, c_vp, c_vp, c_vp, ctypes.POINTER(JindoProof), ctypes.c_double,
                                       ctypes.c_double, ctypes.POINTER(ctypes.c_int)]),
    "rs_jindo_poly_write_to": (ctypes.c_longlong, [c_vp, ctypes.c_int, ctypes.c_int, c_vp, c_size]),
    "rs_jindo_commitment_write_raw": (ctypes.c_longlong, [c_vp, c_vp, c_vp, c_size]),
    "rs_opening_alloc_zero": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp)]),
    "rs_opening_mul_add": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    # ringo_buckler.h
    "rs_buckler_create": (ctypes.c_int, [c_vp, c_vp, ctypes.POINTER(BucklerCircuit), c_vp, c_size, ctypes.POINTER(c_vp)]),
    "rs_buckler_enable_sampling": (ctypes.c_int, [c_vp, ctypes.POINTER(JindoNoise), c_u64p, ctypes.c_int]),
    "rs_buckler_set_device_cosac": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "rs_buckler_destroy": (None, [c_vp]),
    "rs_buckler_batch": (c_size, [c_vp]),
    "rs_buckler_prove": (ctypes.c_int, [c_vp, c_vp, c_vp, ctypes.POINTER(BucklerRand), ctypes.POINTER(BucklerProof), ctypes.POINTER(ctypes.c_double)]),
    "rs_buckler_verify": (ctypes.c_int, [c_vp, c_vp, ctypes.POINTER(BucklerProof), ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_int)]),
    "rs_sha256": (None, [c_vp, c_size, c_vp]),
    "rs_shake128": (None, [c_vp, c_size, c_vp, c_size]),
    # ringo_sampler.h
    "rs_uniform_create": (c_vp, [c_vp]),
    "rs_uniform_destroy": (None, [c_vp]),
    "rs_uniform_sample": (ctypes.c_uint64, [c_vp]),
    "rs_uniform_sample_many": (None, [c_vp, c_vp, c_size]),
    "rs_uniform_sample_n": (ctypes.c_uint64, [c_vp, ctypes.c_uint64]),
    "rs_uniform_sample_float": (ctypes.c_double, [c_vp]),
    "rs_uniform_read": (None, [c_vp, c_vp, c_size]),
    "rs_rounded_create": (c_vp, [c_vp]),
    "rs_rounded_destroy": (None, [c_vp]),
    "rs_rounded_norm_float": (ctypes.c_double, [c_vp]),
    "rs_rounded_sample": (ctypes.c_int64, [c_vp, ctypes.c_double, ctypes.c_double]),
    "rs_cosac_create": (c_vp, [c_vp, c_vp]),
    "rs_cosac_destroy": (None, [c_vp]),
    "rs_cosac_sample": (ctypes.c_int64, [c_vp, ctypes.c_double, ctypes.c_double]),
    "rs_compute_cdt": (ctypes.c_int, [ctypes.c_double, ctypes.c_double, c_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]),
    "rs_tape_builder_create": (ctypes.c_int, [ctypes.POINTER(JindoParams), ctypes.POINTER(JindoNoise), c_u64p, ctypes.c_int, ctypes.c_int,
                                              c_vp, ctypes.POINTER(c_vp)]),
    "rs_tape_builder_create_slots": (ctypes.c_int, [ctypes.POINTER(JindoParams), ctypes.POINTER(JindoNoise), c_u64p, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_int, c_vp, ctypes.POINTER(c_vp)]),
    "rs_tape_builder_set_device_cosac": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "rs_tape_builder_random_elements": (ctypes.c_int, [c_vp, c_vp, c_size]),
    "rs_tape_builder_sum_check_mask": (ctypes.c_int, [c_vp, c_size, c_size, c_vp, c_vp]),
    "rs_tape_builder_destroy": (None, [c_vp]),
    "rs_tape_builder_set_random_bytes": (ctypes.c_int, [c_vp, c_vp, c_size]),
    "rs_tape_builder_start": (ctypes.c_int, [c_vp, c_vp, c_size, ctypes.c_int]),
    "rs_tape_builder_wait": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(JindoTape), ctypes.POINTER(ctypes.c_double)]),
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

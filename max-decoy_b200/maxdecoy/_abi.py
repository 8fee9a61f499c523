"""ctypes view of include/maxdecoy.h (the C ABI of the identification hot path).

The structures mirror the header field by field; `bind()` attaches argtypes/restype to every
exported symbol so a missing symbol fails at load time, not at first use.
"""
import ctypes as C
import os

ALPHABET = "ARNDCEQGHJKMFPOSTUVWY"   # amino_acid.rs:4-5
MAX_PEPTIDE_LEN = 60
WATER_UDA = 18010565
PROTON_UDA = 1007276

MD_OK = 0
STATUS_NAMES = {0: "MD_OK", -1: "MD_ERR_INVALID", -2: "MD_ERR_STATE", -3: "MD_ERR_DEVICE",
                -4: "MD_ERR_NOMEM", -5: "MD_ERR_UNSUPPORTED"}

DECOY_REFERENCE_RANDOM = 0
DECOY_EXHAUSTIVE = 1
DECOY_PERMUTE_TARGET = 2
VARMOD_REFERENCE = 0   # md_varmod_mode
VARMOD_EXPANDED = 1
COMM_ID_BYTES = 128
DECOY_STORED = 0xFFFFFFFF   # `attempt` of a decoy taken from the store (md_decoy_store_set)

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
i16p = C.POINTER(C.c_int16)
f64p = C.POINTER(C.c_double)
f32p = C.POINTER(C.c_float)


class md_config(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_threads", C.c_uint32)]


class md_modification(C.Structure):
    _fields_ = [("accession", C.c_char * 24), ("name", C.c_char * 40), ("position", C.c_uint8),
                ("is_fix", C.c_uint8), ("amino_acid", C.c_uint8), ("_pad", C.c_uint8 * 5),
                ("mono_mass", C.c_int64)]


class md_digest_params(C.Structure):
    _fields_ = [("max_missed_cleavages", C.c_uint32), ("min_len", C.c_uint32), ("max_len", C.c_uint32)]


class md_peptide_table(C.Structure):
    _fields_ = [("n", C.c_uint64), ("seq_bytes", C.c_uint64), ("n_assoc", C.c_uint64),
                ("seq", u8p), ("seq_off", u64p), ("missed_cleavages", u8p), ("weight", i64p),
                ("counts", i16p), ("assoc_off", u64p), ("assoc_protein", u32p)]


class md_index_stats(C.Structure):
    _fields_ = [("n_peptides", C.c_uint64), ("seq_bytes", C.c_uint64), ("device_bytes", C.c_uint64),
                ("min_key", C.c_int64), ("max_key", C.c_int64)]


class md_precursor(C.Structure):
    _fields_ = [("mass", C.c_int64), ("lo", C.c_int64), ("hi", C.c_int64),
                ("charge", C.c_uint32), ("spectrum_id", C.c_uint32)]


class md_candidate_table(C.Structure):
    _fields_ = [("n_spectra", C.c_uint32), ("n", C.c_uint64), ("off", u64p), ("peptide_id", u64p),
                ("var_mask", u64p), ("mod_weight", i64p)]


class md_decoy_table(C.Structure):
    _fields_ = [("n_spectra", C.c_uint32), ("n", C.c_uint64), ("seq_bytes", C.c_uint64),
                ("off", u64p), ("seq", u8p), ("seq_off", u64p), ("var_mask", u64p),
                ("weight", i64p), ("mod_weight", i64p), ("attempt", u32p)]


class md_spectra(C.Structure):
    _fields_ = [("n", C.c_uint32), ("precursor_mz", C.c_void_p), ("charge", C.c_void_p),
                ("spectrum_id", C.c_void_p), ("peak_off", C.c_void_p), ("peak_mz", C.c_void_p),
                ("peak_intensity", C.c_void_p)]


class md_search_params(C.Structure):
    _fields_ = [("lower_ppm", C.c_int64), ("upper_ppm", C.c_int64), ("abs_lower_uda", C.c_int64),
                ("abs_upper_uda", C.c_int64), ("fragment_tolerance", C.c_double),
                ("n_decoys", C.c_uint32), ("decoy_mode", C.c_int32), ("seed", C.c_uint64),
                ("top_k", C.c_uint32), ("min_peaks", C.c_uint32), ("max_fragment_charge", C.c_uint32),
                ("keep_decoys", C.c_uint32)]


class md_psm(C.Structure):
    _fields_ = [("spectrum_id", C.c_uint32), ("rank", C.c_uint16), ("is_decoy", C.c_uint8),
                ("charge", C.c_uint8), ("candidate", C.c_uint64), ("var_mask", C.c_uint64),
                ("mod_weight", C.c_int64), ("raw_score", C.c_int64), ("score", C.c_float),
                ("n_targets", C.c_uint32), ("n_decoys", C.c_uint32), ("_pad", C.c_uint32)]


assert C.sizeof(md_psm) == 56, C.sizeof(md_psm)


class md_identify_stats(C.Structure):
    _fields_ = [("n_spectra", C.c_uint64), ("n_targets", C.c_uint64), ("n_decoys", C.c_uint64),
                ("n_less_decoys", C.c_uint64), ("n_kernel_launches", C.c_uint64),
                ("ms_lookup", C.c_double), ("ms_decoys", C.c_double), ("ms_score", C.c_double),
                ("ms_total", C.c_double), ("ms_kernel_score", C.c_double), ("ms_kernel_decoy", C.c_double),
                ("n_attempts", C.c_uint64), ("n_pairs", C.c_uint64), ("score_bytes", C.c_uint64),
                ("ms_score_prepare", C.c_double), ("n_score_left", C.c_uint64), ("score_pipelined", C.c_uint32), ("_pad", C.c_uint32)]


# numpy dtype with the exact md_psm layout (for zero-copy views of PSM buffers)
PSM_DTYPE = [("spectrum_id", "<u4"), ("rank", "<u2"), ("is_decoy", "u1"), ("charge", "u1"),
             ("candidate", "<u8"), ("var_mask", "<u8"), ("mod_weight", "<i8"), ("raw_score", "<i8"),
             ("score", "<f4"), ("n_targets", "<u4"), ("n_decoys", "<u4"), ("_pad", "<u4")]

ctx_p = C.c_void_p

# name -> (restype, argtypes): every symbol include/maxdecoy.h declares
SYMBOLS = {
    "md_create": (C.c_int, [C.POINTER(md_config), C.POINTER(ctx_p)]),
    "md_destroy": (None, [ctx_p]),
    "md_last_error": (C.c_char_p, [ctx_p]),
    "md_backend_name": (C.c_char_p, []),
    "md_residue_mass": (C.c_int64, [C.c_uint8]),
    "md_sequence_weight": (C.c_int64, [C.c_char_p, C.c_uint32]),
    "md_precursor_window": (C.c_int, [C.c_double, C.c_uint32, C.c_int64, C.c_int64, i64p, i64p, i64p]),
    "md_set_modifications": (C.c_int, [ctx_p, C.POINTER(md_modification), C.c_uint32, C.c_uint32]),
    "md_set_variable_mode": (C.c_int, [ctx_p, C.c_int]),
    "md_substitution_map": (C.c_int, [ctx_p, i64p]),
    "md_digest": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(md_digest_params), u64p]),
    "md_peptides_export": (C.c_int, [ctx_p, C.POINTER(md_peptide_table)]),
    "md_peptide_table_free": (None, [C.POINTER(md_peptide_table)]),
    "md_index_build": (C.c_int, [ctx_p]),
    "md_index_stats_get": (C.c_int, [ctx_p, C.POINTER(md_index_stats)]),
    "md_window_search": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "md_index_export": (C.c_int, [ctx_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "md_candidates": (C.c_int, [ctx_p, C.POINTER(md_precursor), C.c_uint32, C.POINTER(md_candidate_table)]),
    "md_candidate_table_free": (None, [C.POINTER(md_candidate_table)]),
    "md_decoy_store_set": (C.c_int, [ctx_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "md_generate_decoys": (C.c_int, [ctx_p, C.POINTER(md_precursor), C.c_uint32, C.c_uint32, C.c_int,
                                     C.c_uint64, C.POINTER(md_decoy_table)]),
    "md_decoy_table_free": (None, [C.POINTER(md_decoy_table)]),
    "md_identify": (C.c_int, [ctx_p, C.POINTER(md_spectra), C.POINTER(md_search_params), C.c_void_p,
                              C.POINTER(md_identify_stats), C.POINTER(i64p), C.POINTER(u64p)]),
    "md_identify_device": (C.c_int, [ctx_p, C.POINTER(md_spectra), C.POINTER(md_search_params), C.c_void_p,
                                     C.POINTER(md_identify_stats)]),
    "md_comm_unique_id": (C.c_int, [C.c_void_p]),
    "md_comm_init": (C.c_int, [ctx_p, C.c_int32, C.c_int32, C.c_void_p]),
    "md_gather_psms": (C.c_int, [ctx_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "md_comm_destroy": (C.c_int, [ctx_p]),
    "md_sync": (C.c_int, [ctx_p]),
    "md_stream_handle": (C.c_void_p, [ctx_p]),
    "md_last_decoys_export": (C.c_int, [ctx_p, C.POINTER(md_decoy_table)]),
    "md_free": (None, [C.c_void_p]),
}


def bind(path):
    """dlopen `path` and attach the prototypes of every symbol of the header."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library misses a declared symbol
        fn.restype = res
        fn.argtypes = args
    lib._md_path = path
    return lib

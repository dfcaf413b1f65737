"""ctypes binding of the C ABI in include/tagg.h (libtagg.so).

This is the same set of calls a Rust `tagg-sys` crate would make.  There is
no CPU fallback: if libtagg.so is missing the import of any compute entry point fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtagg.so")

# ---- enums (include/tagg.h) ----------------------------------------------------------
OK, ERR_BAD_ARG, ERR_BAD_PLAN, ERR_NO_SUCH_COLUMN, ERR_CUDA, ERR_NCCL, ERR_OOM, ERR_UNSUPPORTED, ERR_NO_DEVICE = range(9)
U64, I64, F64, DATE = range(4)
(OP_TUPLE, OP_COUNT, OP_SUM, OP_MIN, OP_MAX, OP_PERCENTILES, OP_TERMS, OP_HISTOGRAM, OP_FILTER,
 OP_POST_FILTER) = range(10)
PRED_NONE, PRED_RANGE, PRED_LUT = range(3)
DOCSET_ALL, DOCSET_BITSET, DOCSET_SORTED_IDS, DOCSET_COLUMN_RANGE, DOCSET_DEVICE_BITSET = range(5)
ROOT_SCOPE = 0xFFFFFFFF
UNIQUE_ID_BYTES = 128
PATH_AUTO, PATH_GENERIC, PATH_STREAM = 0, 1, 2
READOUT_EAGER, READOUT_LAZY = 0, 1


class Node(C.Structure):
    _fields_ = [("op", C.c_uint8), ("kind", C.c_uint8), ("multi", C.c_uint8), ("pred", C.c_uint8),
                ("field_id", C.c_uint32), ("n_children", C.c_uint32), ("aux", C.c_uint32),
                ("f0", C.c_double), ("f1", C.c_double), ("u0", C.c_uint64), ("u1", C.c_uint64)]


class Blob(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t)]


class Docset(C.Structure):
    _fields_ = [("kind", C.c_int32), ("field_id", C.c_uint32), ("data", C.c_void_p), ("n", C.c_uint64),
                ("lo", C.c_uint64), ("hi", C.c_uint64)]


class FastField(C.Structure):
    _fields_ = [("field_id", C.c_uint32), ("kind", C.c_int32), ("multi", C.c_int32)]


class SegmentInput(C.Structure):
    _fields_ = [("segment", C.c_void_p), ("docset", Docset), ("filters", C.POINTER(Docset)),
                ("n_filters", C.c_uint32)]


class TaggError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"tagg status {status}: {message}")
        self.status = status
        self.message = message


class FastFieldNotAvailableError(TaggError):
    """Mirror of tantivy's FastFieldNotAvailableError (reference sum.rs:50-55, terms.rs:76-81)."""


_lib = None

# name -> (restype, argtypes); every symbol include/tagg.h declares
_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_U64P = C.POINTER(C.c_uint64)
SYMBOLS = {
    "tagg_abi_version": (C.c_uint32, []),
    "tagg_last_error": (C.c_char_p, []),
    "tagg_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "tagg_ctx_create": (C.c_int, [C.c_int, _PP]),
    "tagg_ctx_destroy": (C.c_int, [_P]),
    "tagg_ctx_device": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "tagg_ctx_synchronize": (C.c_int, [_P]),
    "tagg_ctx_set_path": (C.c_int, [_P, C.c_int]),
    "tagg_ctx_timer_start": (C.c_int, [_P]),
    "tagg_ctx_timer_stop": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "tagg_docset_cache": (C.c_int, [_P, C.POINTER(Docset), C.POINTER(Docset)]),
    "tagg_docset_uncache": (C.c_int, [_P, C.POINTER(Docset)]),
    "tagg_docset_to_bitset": (C.c_int, [_P, C.POINTER(Docset), _P, C.c_size_t]),
    "tagg_ctx_launch_count": (C.c_int, [_P, _U64P]),
    "tagg_segment_create": (C.c_int, [_P, C.c_uint32, _PP]),
    "tagg_segment_destroy": (C.c_int, [_P]),
    "tagg_segment_max_doc": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "tagg_column_upload": (C.c_int, [_P, C.c_uint32, C.c_int, _P, C.c_size_t]),
    "tagg_column_upload_codes": (C.c_int, [_P, C.c_uint32, C.c_int, _P, C.c_size_t]),
    "tagg_segment_doc_address_column": (C.c_int, [_P, C.c_uint32, C.c_uint64]),
    "tagg_multicolumn_upload": (C.c_int, [_P, C.c_uint32, C.c_int, _P, C.c_size_t, _P, C.c_size_t]),
    "tagg_multicolumn_upload_codes": (C.c_int, [_P, C.c_uint32, C.c_int, _P, C.c_size_t, _P, C.c_size_t]),
    "tagg_segment_set_deletes": (C.c_int, [_P, _P, C.c_size_t]),
    "tagg_segment_load_fast_file": (C.c_int, [_P, _P, C.c_size_t, _P, C.c_uint32]),
    "tagg_fast_file_entries": (C.c_int, [_P, C.c_size_t, _P, _P, _P, _P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "tagg_column_info": (C.c_int, [_P, C.c_uint32, C.c_int, _U64P, _U64P, C.POINTER(C.c_uint32), _U64P, _U64P]),
    "tagg_column_download": (C.c_int, [_P, C.c_uint32, C.c_int, _P, C.c_size_t]),
    "tagg_plan_create": (C.c_int, [_P, C.POINTER(Node), C.c_uint32, C.POINTER(Blob), C.c_uint32, _PP]),
    "tagg_plan_destroy": (C.c_int, [_P]),
    "tagg_execute": (C.c_int, [_P, C.POINTER(SegmentInput), C.c_uint32, _PP]),
    "tagg_result_free": (C.c_int, [_P]),
    "tagg_execute_begin": (C.c_int, [_P, C.POINTER(SegmentInput), C.c_uint32, _PP]),
    "tagg_pending_wait": (C.c_int, [_P, _PP]),
    "tagg_result_merge": (C.c_int, [_P, _P]),
    "tagg_comm_unique_id": (C.c_int, [_P]),
    "tagg_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "tagg_comm_destroy": (C.c_int, [_P]),
    "tagg_execute_collective": (C.c_int, [_P, C.POINTER(SegmentInput), C.c_uint32, _PP]),
    "tagg_execute_reduce": (C.c_int, [_P, C.POINTER(SegmentInput), C.c_uint32, C.c_int, _PP]),
    "tagg_result_is_local": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "tagg_result_scope_view": (C.c_int, [_P, C.c_uint32, _PP, _PP, _U64P]),
    "tagg_result_metric_view": (C.c_int, [_P, C.c_uint32, _PP, _PP, _U64P]),
    "tagg_plan_set_readout": (C.c_int, [_P, C.c_int]),
    "tagg_result_top_k": (C.c_int, [_P, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, _P, _U64P]),
    "tagg_result_scope_rows": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64, _P, _P]),
    "tagg_result_metric_rows": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64, _P, _P]),
    "tagg_result_scope_len": (C.c_int, [_P, C.c_uint32, _U64P]),
    "tagg_result_scope_read": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint64]),
    "tagg_result_metric_len": (C.c_int, [_P, C.c_uint32, _U64P]),
    "tagg_result_metric_read": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint64]),
    "tagg_result_percentiles_len": (C.c_int, [_P, C.c_uint32, C.c_uint64, _U64P, _U64P]),
    "tagg_result_percentiles_read": (C.c_int, [_P, C.c_uint32, C.c_uint64, _P, _P, C.c_uint64]),
    "tagg_result_stats": (C.c_int, [_P, C.POINTER(C.c_double), _U64P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    # include/tagg_synth.h
    "tagg_synth_column": (C.c_int, [_P, C.c_uint32, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                    C.c_uint64, C.c_uint64, C.c_uint64]),
    "tagg_synth_multicolumn": (C.c_int, [_P, C.c_uint32, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                         C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
}


def lib():
    """Load libtagg.so (once).  Raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the aggregation hot path)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status):
    if status != OK:
        msg = lib().tagg_last_error().decode("utf-8", "replace")
        if status == ERR_NO_SUCH_COLUMN:
            raise FastFieldNotAvailableError(status, msg)
        raise TaggError(status, msg)

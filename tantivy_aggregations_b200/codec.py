"""Host-side value <-> code conversions of tantivy's FastValue types (numpy, vectorised).

Restates tantivy@14735ce common::{f64_to_u64,u64_to_f64,i64_to_u64,u64_to_i64} (external to
/root/reference).  Codes are u64 whose unsigned order equals the value order.
The facade needs them to lower comparison predicates to code ranges and to decode results.
"""
import struct

import numpy as np

from ._ffi import DATE, F64, I64, U64

SIGN = np.uint64(1 << 63)


def f64_to_code(v):
    bits = np.asarray(v, dtype=np.float64).view(np.uint64)
    return np.where((bits >> np.uint64(63)) == 0, bits ^ SIGN, ~bits)


def code_to_f64(c):
    c = np.asarray(c, dtype=np.uint64)
    bits = np.where((c >> np.uint64(63)) == 1, c ^ SIGN, ~c)
    return bits.view(np.float64)


def i64_to_code(v):
    return np.asarray(v, dtype=np.int64).view(np.uint64) ^ SIGN


def code_to_i64(c):
    return (np.asarray(c, dtype=np.uint64) ^ SIGN).view(np.int64)


def values_to_codes(kind, values):
    """Column values in their natural type -> u64 codes (numpy array)."""
    if kind == U64:
        return np.ascontiguousarray(values, dtype=np.uint64)
    if kind in (I64, DATE):
        return np.ascontiguousarray(i64_to_code(values))
    if kind == F64:
        return np.ascontiguousarray(f64_to_code(values))
    raise ValueError(f"bad kind {kind}")


def scalar_code(kind, value):
    return int(values_to_codes(kind, np.array([value]))[0])


def bits_to_value(kind, bits):
    """Value bits as returned by tagg_result_metric_read -> Python value."""
    bits = int(bits)
    if kind == U64:
        return bits
    if kind in (I64, DATE):
        return bits - (1 << 64) if bits >> 63 else bits
    return struct.unpack("<d", struct.pack("<Q", bits))[0]


def f64_bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def num_bits(amplitude):
    """tantivy common::compute_num_bits: widths above 56 are stored as 64."""
    b = int(amplitude).bit_length()
    return b if b <= 56 else 64

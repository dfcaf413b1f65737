//! `extern "C"` surface of `libtagg.so` — one declaration per entry point of `include/tagg.h` (TAGG_ABI_VERSION 2).
//!
//! The reference (`anti-social/tantivy-aggregations`) has no FFI of its own; these are the calls its per-segment
//! collector loop (`src/searcher.rs:27-51`) is replaced by.  Every function returns a `tagg_status` (0 = ok); the
//! message of the last failure on the calling thread is `tagg_last_error()`.  Nothing unwinds across this boundary.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

pub const TAGG_ABI_VERSION: u32 = 2;
pub const TAGG_UNIQUE_ID_BYTES: usize = 128;
pub const TAGG_ROOT_SCOPE: u32 = 0xFFFF_FFFF;

// tagg_status
pub const TAGG_OK: c_int = 0;
pub const TAGG_ERR_BAD_ARG: c_int = 1;
pub const TAGG_ERR_BAD_PLAN: c_int = 2;
pub const TAGG_ERR_NO_SUCH_COLUMN: c_int = 3; // -> tantivy FastFieldNotAvailableError (sum.rs:50-55, terms.rs:76-81 ...)
pub const TAGG_ERR_CUDA: c_int = 4;
pub const TAGG_ERR_NCCL: c_int = 5;
pub const TAGG_ERR_OOM: c_int = 6;
pub const TAGG_ERR_UNSUPPORTED: c_int = 7;
pub const TAGG_ERR_NO_DEVICE: c_int = 8;
// tagg_kind
pub const TAGG_U64: u8 = 0;
pub const TAGG_I64: u8 = 1;
pub const TAGG_F64: u8 = 2;
pub const TAGG_DATE: u8 = 3;
// tagg_op
pub const TAGG_OP_TUPLE: u8 = 0;
pub const TAGG_OP_COUNT: u8 = 1;
pub const TAGG_OP_SUM: u8 = 2;
pub const TAGG_OP_MIN: u8 = 3;
pub const TAGG_OP_MAX: u8 = 4;
pub const TAGG_OP_PERCENTILES: u8 = 5;
pub const TAGG_OP_TERMS: u8 = 6;
pub const TAGG_OP_HISTOGRAM: u8 = 7;
pub const TAGG_OP_FILTER: u8 = 8;
pub const TAGG_OP_POST_FILTER: u8 = 9;
// tagg_pred
pub const TAGG_PRED_NONE: u8 = 0;
pub const TAGG_PRED_RANGE: u8 = 1;
pub const TAGG_PRED_LUT: u8 = 2;
// tagg_docset_kind
pub const TAGG_DOCSET_ALL: i32 = 0;
pub const TAGG_DOCSET_BITSET: i32 = 1;
pub const TAGG_DOCSET_SORTED_IDS: i32 = 2;
pub const TAGG_DOCSET_COLUMN_RANGE: i32 = 3;
pub const TAGG_DOCSET_DEVICE_BITSET: i32 = 4;
// tagg_readout
pub const TAGG_READOUT_EAGER: c_int = 0;
pub const TAGG_READOUT_LAZY: c_int = 1;

#[repr(C)] pub struct tagg_ctx { _p: [u8; 0] }
#[repr(C)] pub struct tagg_segment { _p: [u8; 0] }
#[repr(C)] pub struct tagg_plan { _p: [u8; 0] }
#[repr(C)] pub struct tagg_result { _p: [u8; 0] }
#[repr(C)] pub struct tagg_pending { _p: [u8; 0] }

/// One node of the flattened aggregation tree, pre-order (48 bytes, `struct tagg_node`).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct tagg_node {
    pub op: u8,
    pub kind: u8,
    pub multi: u8,
    pub pred: u8,
    pub field_id: u32,
    pub n_children: u32,
    pub aux: u32,
    pub f0: f64,
    pub f1: f64,
    pub u0: u64,
    pub u1: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct tagg_blob {
    pub data: *const u8,
    pub len: usize,
}

/// 40 bytes, `struct tagg_docset`.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct tagg_docset {
    pub kind: i32,
    pub field_id: u32,
    pub data: *const c_void,
    pub n: u64,
    pub lo: u64,
    pub hi: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct tagg_fast_field {
    pub field_id: u32,
    pub kind: i32,
    pub multi: i32,
}

/// 64 bytes, `struct tagg_segment_input`: one unit of `collect_segment` work (searcher.rs:27-51).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct tagg_segment_input {
    pub segment: *const tagg_segment,
    pub docset: tagg_docset,
    pub filters: *const tagg_docset,
    pub n_filters: u32,
}

extern "C" {
    // ---- library
    pub fn tagg_abi_version() -> u32;
    pub fn tagg_last_error() -> *const c_char;
    pub fn tagg_device_count(out: *mut c_int) -> c_int;
    // ---- context
    pub fn tagg_ctx_create(device: c_int, out: *mut *mut tagg_ctx) -> c_int;
    pub fn tagg_ctx_destroy(ctx: *mut tagg_ctx) -> c_int;
    pub fn tagg_ctx_device(ctx: *const tagg_ctx, out: *mut c_int) -> c_int;
    pub fn tagg_ctx_synchronize(ctx: *mut tagg_ctx) -> c_int;
    pub fn tagg_ctx_set_path(ctx: *mut tagg_ctx, path: c_int) -> c_int;
    pub fn tagg_ctx_timer_start(ctx: *mut tagg_ctx) -> c_int;
    pub fn tagg_ctx_timer_stop(ctx: *mut tagg_ctx, ms: *mut c_double) -> c_int;
    pub fn tagg_ctx_launch_count(ctx: *const tagg_ctx, out: *mut u64) -> c_int;
    // ---- segments and fast-field columns
    pub fn tagg_segment_create(ctx: *mut tagg_ctx, max_doc: u32, out: *mut *mut tagg_segment) -> c_int;
    pub fn tagg_segment_destroy(seg: *mut tagg_segment) -> c_int;
    pub fn tagg_segment_max_doc(seg: *const tagg_segment, out: *mut u32) -> c_int;
    pub fn tagg_column_upload(seg: *mut tagg_segment, field_id: u32, kind: c_int, bytes: *const u8, len: usize) -> c_int;
    pub fn tagg_column_upload_codes(seg: *mut tagg_segment, field_id: u32, kind: c_int, codes: *const u64, n: usize) -> c_int;
    /// u64 fast field `base + doc`, generated on the device: the key column of top_hits (include/tagg.h)
    pub fn tagg_segment_doc_address_column(seg: *mut tagg_segment, field_id: u32, base: u64) -> c_int;
    pub fn tagg_multicolumn_upload(seg: *mut tagg_segment, field_id: u32, kind: c_int, idx_bytes: *const u8, idx_len: usize,
                                   vals_bytes: *const u8, vals_len: usize) -> c_int;
    pub fn tagg_multicolumn_upload_codes(seg: *mut tagg_segment, field_id: u32, kind: c_int, offsets: *const u64, n_offsets: usize,
                                         codes: *const u64, n_codes: usize) -> c_int;
    pub fn tagg_segment_set_deletes(seg: *mut tagg_segment, bytes: *const u8, len: usize) -> c_int;
    pub fn tagg_segment_load_fast_file(seg: *mut tagg_segment, bytes: *const u8, len: usize, fields: *const tagg_fast_field, n_fields: u32) -> c_int;
    pub fn tagg_fast_file_entries(bytes: *const u8, len: usize, fields: *mut u32, idxs: *mut u32, begins: *mut u64, ends: *mut u64,
                                  cap: u32, n_out: *mut u32) -> c_int;
    pub fn tagg_column_info(seg: *const tagg_segment, field_id: u32, which: c_int, min_value: *mut u64, amplitude: *mut u64,
                            num_bits: *mut u32, n_values: *mut u64, packed_len: *mut u64) -> c_int;
    pub fn tagg_column_download(seg: *const tagg_segment, field_id: u32, which: c_int, out: *mut u8, cap: usize) -> c_int;
    // ---- device-resident docsets
    pub fn tagg_docset_cache(seg: *mut tagg_segment, inp: *const tagg_docset, out: *mut tagg_docset) -> c_int;
    pub fn tagg_docset_uncache(seg: *mut tagg_segment, cached: *const tagg_docset) -> c_int;
    pub fn tagg_docset_to_bitset(seg: *const tagg_segment, inp: *const tagg_docset, out: *mut u8, cap: usize) -> c_int;
    // ---- plans
    pub fn tagg_plan_create(ctx: *mut tagg_ctx, nodes: *const tagg_node, n_nodes: u32, blobs: *const tagg_blob, n_blobs: u32,
                            out: *mut *mut tagg_plan) -> c_int;
    pub fn tagg_plan_destroy(plan: *mut tagg_plan) -> c_int;
    pub fn tagg_plan_set_readout(plan: *mut tagg_plan, readout: c_int) -> c_int;
    // ---- execution
    pub fn tagg_execute(plan: *const tagg_plan, inputs: *const tagg_segment_input, n_inputs: u32, out: *mut *mut tagg_result) -> c_int;
    pub fn tagg_result_free(res: *mut tagg_result) -> c_int;
    pub fn tagg_execute_begin(plan: *const tagg_plan, inputs: *const tagg_segment_input, n_inputs: u32, out: *mut *mut tagg_pending) -> c_int;
    pub fn tagg_pending_wait(pending: *mut tagg_pending, out: *mut *mut tagg_result) -> c_int;
    pub fn tagg_result_merge(dst: *mut tagg_result, src: *const tagg_result) -> c_int;
    // ---- multi-GPU
    pub fn tagg_comm_unique_id(out: *mut u8) -> c_int;
    pub fn tagg_comm_init(ctx: *mut tagg_ctx, id: *const u8, rank: c_int, n_ranks: c_int) -> c_int;
    pub fn tagg_comm_destroy(ctx: *mut tagg_ctx) -> c_int;
    pub fn tagg_execute_collective(plan: *const tagg_plan, inputs: *const tagg_segment_input, n_inputs: u32,
                                   out: *mut *mut tagg_result) -> c_int;
    pub fn tagg_execute_reduce(plan: *const tagg_plan, inputs: *const tagg_segment_input, n_inputs: u32, root: c_int,
                               out: *mut *mut tagg_result) -> c_int;
    pub fn tagg_result_is_local(res: *const tagg_result, out: *mut c_int) -> c_int;
    // ---- result readers
    pub fn tagg_result_scope_len(res: *const tagg_result, scope_node: u32, n_buckets: *mut u64) -> c_int;
    pub fn tagg_result_scope_read(res: *const tagg_result, scope_node: u32, keys: *mut u64, parents: *mut u32, cap: u64) -> c_int;
    pub fn tagg_result_metric_len(res: *const tagg_result, node: u32, n_buckets: *mut u64) -> c_int;
    pub fn tagg_result_metric_read(res: *const tagg_result, node: u32, values: *mut u64, seen: *mut u8, cap: u64) -> c_int;
    pub fn tagg_result_scope_view(res: *const tagg_result, scope_node: u32, keys: *mut *const u64, parents: *mut *const u32,
                                  n: *mut u64) -> c_int;
    pub fn tagg_result_metric_view(res: *const tagg_result, node: u32, values: *mut *const u64, seen: *mut *const u8,
                                   n: *mut u64) -> c_int;
    pub fn tagg_result_top_k(res: *mut tagg_result, scope_node: u32, parent_bucket: u64, by_node: u32, k: u64,
                             out_buckets: *mut u32, n_out: *mut u64) -> c_int;
    pub fn tagg_result_scope_rows(res: *mut tagg_result, scope_node: u32, buckets: *const u32, n: u64, keys: *mut u64,
                                  parents: *mut u32) -> c_int;
    pub fn tagg_result_metric_rows(res: *mut tagg_result, node: u32, buckets: *const u32, n: u64, values: *mut u64,
                                   seen: *mut u8) -> c_int;
    pub fn tagg_result_percentiles_len(res: *const tagg_result, node: u32, bucket: u64, n_total: *mut u64, n_pairs: *mut u64) -> c_int;
    pub fn tagg_result_percentiles_read(res: *const tagg_result, node: u32, bucket: u64, ranks: *mut u64, value_bits: *mut u64,
                                        cap: u64) -> c_int;
    pub fn tagg_result_stats(res: *const tagg_result, kernel_ms: *mut c_double, alg_bytes: *mut u64, n_launches: *mut u32,
                             path_used: *mut u32) -> c_int;
}

#[cfg(test)]
mod tests {
    use super::*;
    #[test]
    fn struct_layouts_match_the_header() {
        assert_eq!(std::mem::size_of::<tagg_node>(), 48);
        assert_eq!(std::mem::size_of::<tagg_docset>(), 40);
        assert_eq!(std::mem::size_of::<tagg_segment_input>(), 64);
    }
}

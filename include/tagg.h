/*
 * tagg.h — C ABI of the B200-native aggregation hot path (libtagg.so).
 *
 * This is the drop-in boundary for ONE path of anti-social/tantivy-aggregations:
 * the per-segment collector loop `collect_segment` (reference src/searcher.rs:27-51)
 * and every `SegmentAgg::collect` it inlines (src/metric/{count,sum,minmax,percentile}.rs, src/bucket/{terms,histogram}.rs,
 * src/filter.rs, src/post_filter.rs, src/tuple.rs).  The reference has no FFI of its
 * own (pure safe Rust); the entry points below are what a `tagg-sys` crate would bind
 * so that `AggSearcher::agg_search` (src/searcher.rs:12-25) keeps its signature while
 * the per-document loop runs on the GPU.  INTEGRATION.md shows that binding.
 *
 * Conventions
 *   - every function returns a tagg_status (0 = ok); the message of the last failure
 *     on the calling thread is available from tagg_last_error();
 *   - no exception, abort or unwind crosses this boundary;
 *   - host buffers passed in are borrowed for the duration of the call only;
 *   - handles are owned by the library until the matching *_destroy / *_free;
 *   - tagg_ctx / tagg_segment / tagg_plan are immutable after construction and may be
 *     shared between host threads; a tagg_result is confined to one thread
 *     (mirrors `PreparedAgg: Sync`, `Fruit: Send`, reference src/agg.rs:10-28).
 *
 * There is NO CPU fallback: every entry point that computes needs a CUDA device and
 * fails with TAGG_ERR_NO_DEVICE / TAGG_ERR_CUDA otherwise.
 */
#ifndef TAGG_H
#define TAGG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAGG_ABI_VERSION 2

typedef enum tagg_status {
    TAGG_OK = 0,
    TAGG_ERR_BAD_ARG = 1,
    TAGG_ERR_BAD_PLAN = 2,
    TAGG_ERR_NO_SUCH_COLUMN = 3, /* maps to tantivy FastFieldNotAvailableError (sum.rs:50-55, terms.rs:76-81 ...) */
    TAGG_ERR_CUDA = 4,
    TAGG_ERR_NCCL = 5,
    TAGG_ERR_OOM = 6,
    TAGG_ERR_UNSUPPORTED = 7,
    TAGG_ERR_NO_DEVICE = 8
} tagg_status;

/* Value type of a fast-field column (tantivy FastValue: u64, i64, f64, DateTime).
 * Columns always hold order-preserving u64 *codes*:
 *   u64: identity; i64/date: v ^ (1<<63); f64: sign==0 ? bits ^ (1<<63) : ~bits. */
typedef enum tagg_kind { TAGG_U64 = 0, TAGG_I64 = 1, TAGG_F64 = 2, TAGG_DATE = 3 } tagg_kind;

/* Aggregation node opcodes — one per reference constructor family. */
typedef enum tagg_op {
    TAGG_OP_TUPLE = 0,       /* (a1, .., an) fan-out                      tuple.rs:63-67   */
    TAGG_OP_COUNT = 1,       /* count_agg()                               count.rs:53-55   */
    TAGG_OP_SUM = 2,         /* sum_agg_{u64,i64,f64}[s]                  sum.rs:95-102,131-140 */
    TAGG_OP_MIN = 3,         /* min_agg_{u64,i64,f64,date}[s]             minmax.rs:97-106,135-145 */
    TAGG_OP_MAX = 4,         /* max_agg_*                                 minmax.rs (gt arm) */
    TAGG_OP_PERCENTILES = 5, /* percentiles_agg_f64[s]                    percentile.rs:87-90,119-124 */
    TAGG_OP_TERMS = 6,       /* terms_agg_{u64,i64}[s](field, sub)        terms.rs:127-132,172-179 */
    TAGG_OP_HISTOGRAM = 7,   /* histogram_agg_f64(field,start,interval,sub) histogram.rs:136-152 */
    TAGG_OP_FILTER = 8,      /* filter_agg(&query, sub): AND a 2nd docset filter.rs:100-122 */
    TAGG_OP_POST_FILTER = 9  /* post_filter_agg_*(field, pred, sub)       post_filter.rs:245-249,289-297 */
} tagg_op;

/* Declarative predicates a POST_FILTER node can carry (the reference takes a Rust
 * closure; the facade lowers it: comparison closures -> RANGE on codes, arbitrary
 * closures over a small code domain -> LUT, anything else -> the host evaluates the
 * closure per doc into a bitset and uses a FILTER node). */
typedef enum tagg_pred {
    TAGG_PRED_NONE = 0,
    TAGG_PRED_RANGE = 1, /* passes iff u0 <= code <= u1 (unsigned compare on codes)          */
    TAGG_PRED_LUT = 2    /* passes iff code in [u0, u0+u1) and bit (code-u0) of blob[aux] set */
} tagg_pred;

/* One node of the flattened aggregation tree, pre-order: a node is followed
 * immediately by its n_children sub-trees. */
typedef struct tagg_node {
    uint8_t op;          /* tagg_op */
    uint8_t kind;        /* tagg_kind of the column the node reads (ignored for TUPLE/COUNT/FILTER) */
    uint8_t multi;       /* 1 = multi-valued fast field (the *s constructors) */
    uint8_t pred;        /* tagg_pred (POST_FILTER only) */
    uint32_t field_id;   /* column to read (tantivy Field id) */
    uint32_t n_children; /* TUPLE: 2..10; TERMS/HISTOGRAM/FILTER/POST_FILTER: 1; leaves: 0 */
    uint32_t aux;        /* FILTER: index into tagg_segment_input.filters; PRED_LUT: blob index */
    double f0;           /* HISTOGRAM: start    (histogram.rs:9-21) */
    double f1;           /* HISTOGRAM: interval */
    uint64_t u0;         /* PRED_RANGE: lo code; PRED_LUT: base code */
    uint64_t u1;         /* PRED_RANGE: hi code; PRED_LUT: number of bits in the LUT */
} tagg_node;

typedef struct tagg_blob {
    const uint8_t* data; /* LUT bitmap, bit i = data[i>>3] >> (i&7) & 1 */
    size_t len;
} tagg_blob;

typedef enum tagg_docset_kind {
    TAGG_DOCSET_ALL = 0,        /* AllScorer: every doc in 0..max_doc                        */
    TAGG_DOCSET_BITSET = 1,     /* data = ceil(max_doc/8) bytes, doc d set iff data[d>>3]>>(d&7)&1 */
    TAGG_DOCSET_SORTED_IDS = 2, /* data = n strictly ascending uint32 doc ids                */
    TAGG_DOCSET_COLUMN_RANGE = 3,/* docset produced on the device: docs whose single-valued fast
                                    field `field_id` has lo <= code <= hi (TermQuery / RangeQuery on
                                    an INDEXED|FAST field).  data = NULL.          */
    TAGG_DOCSET_DEVICE_BITSET = 4 /* a bitset already resident in HBM: data = the device pointer
                                    returned by tagg_docset_cache (reusable filters stay on the GPU) */
} tagg_docset_kind;

typedef struct tagg_docset {
    int32_t kind;     /* tagg_docset_kind */
    uint32_t field_id;/* COLUMN_RANGE */
    const void* data; /* BITSET: bytes; SORTED_IDS: uint32_t[n] */
    uint64_t n;       /* BITSET: byte length; SORTED_IDS: id count */
    uint64_t lo, hi;  /* COLUMN_RANGE, inclusive, on codes */
} tagg_docset;

typedef struct tagg_ctx tagg_ctx;
typedef struct tagg_segment tagg_segment;
typedef struct tagg_plan tagg_plan;
typedef struct tagg_result tagg_result;

/* One unit of `collect_segment` work (searcher.rs:27-51): a segment, the main query's
 * matched docs for it, and one docset per FILTER node of the plan (filter.rs:65-73). */
typedef struct tagg_segment_input {
    const tagg_segment* segment;
    tagg_docset docset;
    const tagg_docset* filters;
    uint32_t n_filters;
} tagg_segment_input;

/* ---- library ---------------------------------------------------------------------- */
uint32_t tagg_abi_version(void);
const char* tagg_last_error(void);            /* thread-local, never NULL */
int tagg_device_count(int* out);

/* ---- context: one per GPU (one process per GPU when sharded, §8e) ------------------- */
int tagg_ctx_create(int device, tagg_ctx** out);
int tagg_ctx_destroy(tagg_ctx* ctx);
int tagg_ctx_device(const tagg_ctx* ctx, int* out);
int tagg_ctx_synchronize(tagg_ctx* ctx);
/* Tuning/testing knob: 0 = let the planner choose (default), 1 = force the generic
 * tree-walking kernel, 2 = force the streaming kernels (error if the plan has no fast shape). */
int tagg_ctx_set_path(tagg_ctx* ctx, int path);
/* Device-side stopwatch (CUDA events on the stream executes run on): start, run K executes from the
 * same host thread, stop -> elapsed milliseconds on the device timeline. */
int tagg_ctx_timer_start(tagg_ctx* ctx);
int tagg_ctx_timer_stop(tagg_ctx* ctx, double* ms);
/* Number of kernels launched by this context so far (bench.py `gpu_launches`). */
int tagg_ctx_launch_count(const tagg_ctx* ctx, uint64_t* out);

/* ---- segments and fast-field columns (replaces SegmentReader::fast_fields(),
 *      reference call sites sum.rs:50, minmax.rs:50, terms.rs:76, histogram.rs:81,
 *      percentile.rs:49, post_filter.rs:197) ------------------------------------------ */
int tagg_segment_create(tagg_ctx* ctx, uint32_t max_doc, tagg_segment** out);
int tagg_segment_destroy(tagg_segment* seg);
int tagg_segment_max_doc(const tagg_segment* seg, uint32_t* out);

/* Single-valued column from tantivy's own bytes (FastFieldReader payload):
 * min_value:u64 LE | amplitude:u64 LE | LSB-first bit-packed deltas | >=7 pad bytes. */
int tagg_column_upload(tagg_segment* seg, uint32_t field_id, int kind,
                       const uint8_t* bytes, size_t len);
/* Single-valued column from decoded codes (host decoded with FastFieldReader::get);
 * the device computes min/amplitude and re-packs to the same layout. n == max_doc. */
int tagg_column_upload_codes(tagg_segment* seg, uint32_t field_id, int kind,
                             const uint64_t* codes, size_t n);
/* Multi-valued column (MultiValueIntFastFieldReader): idx column of max_doc+1 offsets
 * and vals column of codes, both in the single-valued byte layout above. */
int tagg_multicolumn_upload(tagg_segment* seg, uint32_t field_id, int kind,
                            const uint8_t* idx_bytes, size_t idx_len,
                            const uint8_t* vals_bytes, size_t vals_len);
int tagg_multicolumn_upload_codes(tagg_segment* seg, uint32_t field_id, int kind,
                                  const uint64_t* offsets, size_t n_offsets, /* max_doc+1 */
                                  const uint64_t* codes, size_t n_codes);
/* The address column of a segment: a u64 fast field whose value for document d is base + d, generated on the device.  Not a
 * tantivy fast field: it gives documents a key, so that "the matched documents with the k best values of a field"
 * (top_hits, reference README.md:31-45) is a terms aggregation keyed by it (Searcher.top_hits in the Python mirror). */
int tagg_segment_doc_address_column(tagg_segment* seg, uint32_t field_id, uint64_t base);
/* A whole tantivy `.fast` CompositeFile (the mmap'd segment file): the footer is parsed on the host and every
 * requested field's payload(s) go through tagg_column_upload / tagg_multicolumn_upload unchanged — no host decode.
 * Layout restated from tantivy@14735ce common/composite_file.rs (see csrc/columns.cu); NOT pinned to real tantivy bytes. */
typedef struct tagg_fast_field { uint32_t field_id; int32_t kind; int32_t multi; } tagg_fast_field;
int tagg_segment_load_fast_file(tagg_segment* seg, const uint8_t* bytes, size_t len, const tagg_fast_field* fields, uint32_t n_fields);
/* The directory of such a file (host only, no device needed): (field, idx) and byte range of every payload. */
int tagg_fast_file_entries(const uint8_t* bytes, size_t len, uint32_t* fields, uint32_t* idxs, uint64_t* begins, uint64_t* ends,
                           uint32_t cap, uint32_t* n_out);
/* DeleteBitSet bytes: doc d is DELETED iff bytes[d>>3]>>(d&7)&1 (searcher.rs:41-46). */
int tagg_segment_set_deletes(tagg_segment* seg, const uint8_t* bytes, size_t len);
/* Introspection (tests): header and packed bytes of a resident column.
 * which: 0 = single-valued column / vals column of a multi field, 1 = idx column. */
int tagg_column_info(const tagg_segment* seg, uint32_t field_id, int which,
                     uint64_t* min_value, uint64_t* amplitude, uint32_t* num_bits,
                     uint64_t* n_values, uint64_t* packed_len);
int tagg_column_download(const tagg_segment* seg, uint32_t field_id, int which,
                         uint8_t* out, size_t cap);

/* ---- device-resident docsets (SURVEY §7 "PCIe handoff", §8f-1) ------------------------------
 * tagg_docset_cache evaluates / uploads `in` once into a bitset owned by the segment and fills
 * `out` with a TAGG_DOCSET_DEVICE_BITSET docset that later executes read straight from HBM (a
 * reusable filter such as status=0).  Released by tagg_docset_uncache or with the segment.
 * tagg_docset_to_bitset evaluates a docset on the device and returns the bitset bytes to the host
 * (ceil(max_doc/8) bytes) — e.g. a TermQuery on an INDEXED|FAST field without touching postings. */
int tagg_docset_cache(tagg_segment* seg, const tagg_docset* in, tagg_docset* out);
int tagg_docset_uncache(tagg_segment* seg, const tagg_docset* cached);
int tagg_docset_to_bitset(const tagg_segment* seg, const tagg_docset* in, uint8_t* out, size_t cap);

/* ---- plans (replaces Agg::prepare -> PreparedAgg, agg.rs:10-28) ---------------------- */
int tagg_plan_create(tagg_ctx* ctx, const tagg_node* nodes, uint32_t n_nodes,
                     const tagg_blob* blobs, uint32_t n_blobs, tagg_plan** out);
int tagg_plan_destroy(tagg_plan* plan);
/* How a result of this plan is read out.  EAGER (default): the compact fruit image is downloaded with the execute call.
 * LAZY: only the bucket counts are; the image stays in HBM and tagg_result_top_k / tagg_result_*_rows fetch just the rows
 * they need (the plain readers still work: the first one downloads everything). */
typedef enum tagg_readout { TAGG_READOUT_EAGER = 0, TAGG_READOUT_LAZY = 1 } tagg_readout;
int tagg_plan_set_readout(tagg_plan* plan, int readout);

/* ---- execution (replaces collect_segment, searcher.rs:27-51) ------------------------
 * All inputs are folded into ONE fruit, like Executor::SingleThread threading one
 * harvest through every segment (searcher.rs:66-78).  Call once per segment and
 * tagg_result_merge() the results for the Executor::ThreadPool shape (:79-98). */
int tagg_execute(const tagg_plan* plan, const tagg_segment_input* inputs, uint32_t n_inputs,
                 tagg_result** out);
int tagg_result_free(tagg_result* res);
/* The same call split at its one synchronisation point, so that a host can keep two (or more) queries in flight — the
 * preparation of query i+1 then overlaps the kernels of query i (the reference overlaps segments on its thread pool,
 * searcher.rs:79-92; here the unit in flight is the whole query).  tagg_execute_begin returns as soon as the pass, the
 * compaction and the download are queued; host buffers handed in (docsets) stay borrowed until tagg_pending_wait returns,
 * which always consumes the pending handle. */
typedef struct tagg_pending tagg_pending;
int tagg_execute_begin(const tagg_plan* plan, const tagg_segment_input* inputs, uint32_t n_inputs, tagg_pending** out);
int tagg_pending_wait(tagg_pending* pending, tagg_result** out);
/* PreparedAgg::merge (count.rs:39-41, sum.rs:59-70, minmax.rs:59-72, terms.rs:85-92,
 * histogram.rs:90-97); both results must come from the same plan. */
int tagg_result_merge(tagg_result* dst, const tagg_result* src);

/* ---- multi-GPU: one process per GPU; bucket tables merged over NCCL (§8e) ------------ */
#define TAGG_UNIQUE_ID_BYTES 128
int tagg_comm_unique_id(uint8_t out[TAGG_UNIQUE_ID_BYTES]);      /* rank 0, then broadcast by the host */
int tagg_comm_init(tagg_ctx* ctx, const uint8_t id[TAGG_UNIQUE_ID_BYTES], int rank, int n_ranks);
int tagg_comm_destroy(tagg_ctx* ctx);
/* Like tagg_execute over this rank's segments, then one collective merge step;
 * every rank receives the merged fruit.  Collective: all ranks must call it, with plans of the same tree.
 * Every call agrees the bucket-table layout across ranks with one tiny all-reduce (ranks may change their segment
 * sets between calls independently); the previous agreement is used optimistically while that all-reduce is in flight. */
int tagg_execute_collective(const tagg_plan* plan, const tagg_segment_input* inputs,
                            uint32_t n_inputs, tagg_result** out);
/* The same, but the bucket tables are merged by an NCCL reduce into rank `root` only (the host that answers the query,
 * searcher.rs:93-96 has ONE harvest): on the root *out is the merged fruit; on the other ranks it is an empty placeholder
 * (tagg_result_is_local -> 0, every scope / metric length 0).  Plans whose tables cannot be reduced cell by cell (hashed
 * scopes, percentiles, exact NaN / signed-zero f64 min / max) exchange compact results instead and every rank gets the fruit. */
int tagg_execute_reduce(const tagg_plan* plan, const tagg_segment_input* inputs,
                        uint32_t n_inputs, int root, tagg_result** out);
int tagg_result_is_local(const tagg_result* res, int* out);

/* ---- result readers -----------------------------------------------------------------
 * A plan's bucket scopes are: the root (scope_node = UINT32_MAX, exactly one bucket) and
 * one per TERMS / HISTOGRAM node (scope_node = that node's index).  Buckets of a scope are
 * returned in a stable order; `parents[i]` is the index of bucket i's enclosing bucket in
 * the parent scope's order.  Only buckets that at least one document reached exist
 * (`entry().or_insert_with`, terms.rs:129-130, histogram.rs:148-149). */
#define TAGG_ROOT_SCOPE 0xFFFFFFFFu
int tagg_result_scope_len(const tagg_result* res, uint32_t scope_node, uint64_t* n_buckets);
/* keys: TERMS -> key value bits (u64, or i64 two's complement); HISTOGRAM -> bucket_ord. */
int tagg_result_scope_read(const tagg_result* res, uint32_t scope_node,
                           uint64_t* keys, uint32_t* parents, uint64_t cap);
/* Leaf metric `node` for every bucket of its enclosing scope, in that scope's order.
 * values: COUNT -> count; SUM/MIN/MAX -> value bits in the column's type (u64; i64/date
 * two's complement; f64 IEEE bits).  seen[i] == 0 <=> the reference's Option is None. */
int tagg_result_metric_len(const tagg_result* res, uint32_t node, uint64_t* n_buckets);
int tagg_result_metric_read(const tagg_result* res, uint32_t node,
                            uint64_t* values, uint8_t* seen, uint64_t cap);
/* Zero-copy views of the same arrays: pointers into the result's (page-locked) image, valid until the result is freed
 * or merged into.  n = number of buckets. */
int tagg_result_scope_view(const tagg_result* res, uint32_t scope_node,
                           const uint64_t** keys, const uint32_t** parents, uint64_t* n);
int tagg_result_metric_view(const tagg_result* res, uint32_t node,
                            const uint64_t** values, const uint8_t** seen, uint64_t* n);
/* Terms::top_k(k, |bucket| <leaf metric by_node>) (terms.rs:425-457) on the device-resident image: the (at most) k buckets
 * of the TERMS scope `scope_node` under parent bucket `parent_bucket` (0 for a top-level terms_agg) with the largest value
 * of the leaf `by_node`, descending; equal values ascending by key (the reference's heap order).  Option metrics order as
 * Rust's Option (None below every Some); f64 by the total order of the codes.  out_buckets: bucket indices in the order of
 * tagg_result_scope_read; *n_out of them are written.  A radix select on the GPU when the image is there, else on the host. */
int tagg_result_top_k(tagg_result* res, uint32_t scope_node, uint64_t parent_bucket, uint32_t by_node, uint64_t k,
                      uint32_t* out_buckets, uint64_t* n_out);
/* Rows by bucket index (e.g. the winners of tagg_result_top_k): gathered on the device for a lazily read result. */
int tagg_result_scope_rows(tagg_result* res, uint32_t scope_node, const uint32_t* buckets, uint64_t n,
                           uint64_t* keys, uint32_t* parents);
int tagg_result_metric_rows(tagg_result* res, uint32_t node, const uint32_t* buckets, uint64_t n,
                            uint64_t* values, uint8_t* seen);
/* PERCENTILES leaf (root scope or nested): an exact rank summary for bucket `bucket`.
 * n_total = values inserted; pairs (ranks[i], value_bits[i]) are exact 1-based order
 * statistics, ascending.  percentile(q) picks the pair nearest the CKMS target rank
 * (percentile.rs:163-165). */
int tagg_result_percentiles_len(const tagg_result* res, uint32_t node, uint64_t bucket,
                                uint64_t* n_total, uint64_t* n_pairs);
int tagg_result_percentiles_read(const tagg_result* res, uint32_t node, uint64_t bucket,
                                 uint64_t* ranks, uint64_t* value_bits, uint64_t cap);
/* Timing of the last execute that produced `res`: device time (CUDA events, ms) of the
 * kernels only, and the algorithmic bytes they streamed (SURVEY §8d B_alg). */
int tagg_result_stats(const tagg_result* res, double* kernel_ms, uint64_t* alg_bytes,
                      uint32_t* n_launches, uint32_t* path_used);

#ifdef __cplusplus
}
#endif
#endif /* TAGG_H */

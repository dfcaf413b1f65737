// host.h — host-side internals of libtagg.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/tagg.h"
#include "../../include/tagg_synth.h"
#include "dev.cuh"

// ---- error plumbing ---------------------------------------------------------------------------
int tagg_fail(int status, const char* fmt, ...);
#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return tagg_fail(_e == cudaErrorMemoryAllocation ? TAGG_ERR_OOM : TAGG_ERR_CUDA,   \
                             "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---- handles ------------------------------------------------------------------------------------
// Per-call host resources, pooled in the context: a stream, two events and a pinned staging block
// (small control uploads and small result downloads go through pinned memory so they are truly async).
struct CallRes {
    // Device scratch of the calls that run on this slot (arena, descriptors, ...), kept between calls: a block is only
    // ever touched on THIS slot's stream.  (Returning it to the stream-ordered pool instead makes the next call in flight —
    // on another stream — pick the block up with a dependency on this call's tail: two queries in flight then serialise.)
    struct DevBlock { void* p; size_t bytes; bool in_use; };
    std::vector<DevBlock> blocks;
    cudaStream_t st = nullptr;   // uploads, downloads, ordering
    cudaStream_t st2 = nullptr;  // upload stream: host docsets cross PCIe here while kernels of earlier chunks run on st
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t chain_ev = nullptr;  // this call's pass AND compaction are done (the next call in flight starts its pass behind it)
    cudaEvent_t chunk_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, join_ev = nullptr;
    uint8_t* pinned = nullptr;
    size_t pinned_bytes = 0, pinned_used = 0;
};

struct tagg_ctx {
    int device = 0;
    int sm_count = 148;
    int path = 0;
    std::atomic<uint64_t> launches{0};
    std::mutex mu;
    std::vector<cudaStream_t> stream_pool;
    cudaEvent_t timer0 = nullptr, timer1 = nullptr;
    // With several calls in flight (tagg_execute_begin) the passes still run one after the other — a persistent pass owns
    // every SM (its registers leave no room for another block) — so each call's pass explicitly waits for the previous
    // call's pass AND compaction: its CUDA-event bracket then measures the pass itself and not the queueing behind another
    // query, and the previous call's small compaction kernels are not starved until this pass ends (their download then
    // overlaps this pass).  Guarded by mu.
    cudaEvent_t last_pass_done = nullptr;
    // NCCL (comm.cu), loaded lazily with dlopen so that single-GPU use has no NCCL dependency
    void* nccl = nullptr;  // opaque NcclState*
    int rank = 0, n_ranks = 1;

    uint64_t* pct_sched_dev = nullptr;  // pct.cu: the geometric rank schedule of the exact lists (uploaded once)
    uint32_t pct_sched_len = 0;

    std::vector<CallRes*> call_pool;
    // freed results are recycled: their arrays keep their capacity, so the next result of the same shape is filled without
    // fresh allocations (a 100 k-bucket result is ~4 MB of vectors: page faults and zero fill were ~40 % of its readout)
    std::vector<struct tagg_result*> result_pool;

    cudaStream_t acquire_stream();
    void release_stream(cudaStream_t s);
    CallRes* acquire_call();
    void release_call(CallRes* c);
};

struct HostColumn {
    void* dptr = nullptr;       // allocation (payload, zero padded)
    size_t alloc_bytes = 0;
    size_t payload_bytes = 0;   // ceil(n_values * num_bits / 8)
    uint64_t min_value = 0, amplitude = 0, n_values = 0;
    uint32_t num_bits = 0;
    int kind = 0;
    DevColumn dev() const {
        DevColumn d;
        d.words = (const uint64_t*)dptr;
        d.min_value = min_value;
        d.max_value = min_value + amplitude;
        d.mask = num_bits == 64 ? ~0ull : ((1ull << num_bits) - 1ull);
        d.n_values = n_values;
        d.num_bits = num_bits;
        d.kind = (uint32_t)kind;
        return d;
    }
};

struct tagg_segment {
    tagg_ctx* ctx = nullptr;
    uint32_t max_doc = 0;
    std::unordered_map<uint32_t, HostColumn> cols;                          // single-valued
    std::unordered_map<uint32_t, std::pair<HostColumn, HostColumn>> mcols;  // (idx, vals)
    uint32_t* d_deleted = nullptr;
    bool has_deletes = false;
    uint64_t n_deleted = 0;
    std::vector<uint32_t*> cached_bitsets;  // tagg_docset_cache allocations
    std::mutex mu;
};

// Derived, immutable description of a plan (shared with its results).
struct PlanMeta {
    std::vector<tagg_node> nodes;
    std::vector<uint16_t> end;        // one past the sub-tree
    std::vector<int> parent_node;     // -1 for the root node
    std::vector<int> scope_of;        // enclosing scope of each node
    std::vector<int> own_scope;       // TERMS / HISTOGRAM: scope keyed by the node, else -1
    std::vector<int> slot_of;         // leaf metric: slot id, else -1
    std::vector<int> pct_of;          // PERCENTILES: percentile slot id, else -1
    std::vector<int> scope_node;      // scope -> node index (-1 for root)
    std::vector<int> scope_parent;    // scope -> parent scope (-1 for root)
    std::vector<int> slot_node;       // slot -> node
    std::vector<int> pct_node;
    struct ColRef { uint32_t field_id; int multi; };
    std::vector<ColRef> colrefs;      // device column slots: single -> 1 slot, multi -> 2 (idx, vals)
    std::vector<int> col_slot;        // node -> first device column slot, -1 if none
    uint32_t n_filters = 0;
    std::vector<std::vector<uint8_t>> blobs;
};

// percentiles on the streaming path (pct.cu): the rank-bin thresholds a sample pass found for (plan, segment set).  They
// only steer efficiency — every pass verifies its own precision on the real counts — so repeated queries reuse them and
// skip the sample kernel, its sort and the host round trip.
struct PctThresholds {
    bool valid = false;
    std::vector<const void*> segs;
    uint32_t node = 0;
    uint64_t lo = 0, span = 0;
    uint32_t shift = 0, mul = 0, n_bins = 0;
    bool linear = false;
    double f_lo = 0, f_scale = 0;
    uint64_t tail_seen = 0;  // values that went to the exact lists last time (sizes the next list)
};

struct tagg_plan {
    tagg_ctx* ctx = nullptr;
    std::shared_ptr<PlanMeta> meta;
    mutable PctThresholds pct_cache[4];  // guarded by mu
    // k_mterms' shared-memory hot-key front, per TERMS node: 0 = measure, 1 = it found reuse, 2.. = it did not (counts the
    // queries since, mterms.cu); guarded by mu
    mutable uint8_t mt_front_hint[64] = {};  // TAGG_MAX_NODES (dev.cuh)
    int readout = 0;                     // TAGG_READOUT_*
    std::vector<uint8_t*> d_blobs;  // device copies of LUT bitmaps
    // multi-GPU: the key domains agreed across ranks on the previous collective call of this plan; reused optimistically
    // and re-verified by every call's own agreement (exec.cu)
    mutable std::mutex mu;
    mutable std::vector<const void*> dom_key;
    mutable std::vector<uint64_t> dom_local, dom_agreed;
};

struct PctSummary {
    uint64_t n_total = 0;
    std::vector<uint64_t> ranks, value_bits;       // exact (rank, value) pairs, ascending
    std::vector<uint64_t> rank_lo, rank_hi;        // after a merge ranks are intervals; empty = exact
};

struct tagg_result {
    tagg_ctx* ctx = nullptr;
    std::shared_ptr<PlanMeta> meta;
    // Two representations of the bucket scopes and leaf metrics:
    //  (a) has_img: the compact image produced on the device (compact.cu), downloaded into `img` (page-locked, recycled
    //      with the result): [buckets per scope u64][per scope: keys u64, parents u32][per slot: values u64, seen u8] at
    //      the byte offsets below.  The readers copy straight out of it.
    //  (b) the vectors: after a host-side merge (PreparedAgg::merge on compact results, result.cu); materialize() moves
    //      (a) into (b).
    uint8_t* img = nullptr;
    size_t img_cap = 0;
    bool img_pinned = false, has_img = false;
    std::vector<uint64_t> n_scope;
    std::vector<size_t> off_keys, off_parents, off_values, off_seen;
    // the same image in HBM (arrays at capacity stride), kept for top_k / row reads on the device
    uint8_t* d_img = nullptr;
    size_t d_bytes = 0;
    cudaStream_t d_stream = nullptr;
    std::vector<size_t> d_off_keys, d_off_parents, d_off_values, d_off_seen;
    bool lazy = false;  // arrays were not downloaded yet (TAGG_READOUT_LAZY): result_ensure_host fetches them on demand
    struct Scope { std::vector<uint64_t> keys; std::vector<uint32_t> parents; };
    struct Slot { std::vector<uint64_t> values; std::vector<uint8_t> seen; };
    std::vector<Scope> scopes;   // by scope id
    std::vector<Slot> slots;     // by slot id
    std::vector<std::unordered_map<uint64_t, PctSummary>> pcts;  // by pct slot: bucket -> summary
    double kernel_ms = 0;
    uint64_t alg_bytes = 0;
    uint32_t n_launches = 0;
    uint32_t path_used = 0;
    uint32_t merged_elsewhere = 0;  // tagg_execute_reduce on a non-root rank: the fruit lives on the root

    uint64_t scope_len(size_t s) const { return has_img ? n_scope[s] : scopes[s].keys.size(); }
    const uint64_t* scope_keys(size_t s) const { return has_img ? (const uint64_t*)(img + off_keys[s]) : scopes[s].keys.data(); }
    const uint32_t* scope_parents(size_t s) const { return has_img ? (const uint32_t*)(img + off_parents[s]) : scopes[s].parents.data(); }
    uint64_t slot_len(size_t k) const { return has_img ? n_scope[meta->scope_of[meta->slot_node[k]]] : slots[k].values.size(); }
    const uint64_t* slot_values(size_t k) const { return has_img ? (const uint64_t*)(img + off_values[k]) : slots[k].values.data(); }
    const uint8_t* slot_seen(size_t k) const { return has_img ? img + off_seen[k] : slots[k].seen.data(); }
    void materialize();     // (a) -> (b)
    void release_device();  // frees d_img
    ~tagg_result();
};

// ---- kernels' launchers (generic.cu, stream.cu, columns.cu) ------------------------------------------
cudaError_t launch_generic(const DevPlan* dplan, const DevSegment* dseg, uint64_t n_cand, uint32_t seg_index, int sm_count,
                           cudaStream_t stream);
cudaError_t launch_edge_fixup(const DevSegment* segs, int col, int is_min, uint64_t* acc, const uint8_t* seen, const uint64_t* edge,
                              uint64_t cap, int sm_count, cudaStream_t stream);

// columns.cu
int column_from_bytes(tagg_ctx* ctx, int kind, const uint8_t* bytes, size_t len, uint64_t n_values, HostColumn* out);
int column_from_device_codes(tagg_ctx* ctx, int kind, const uint64_t* d_codes, uint64_t n, HostColumn* out,
                             cudaStream_t stream);
void column_free(HostColumn* c);
cudaError_t launch_ids_to_bitset(const uint32_t* ids, uint64_t n, uint32_t* words, uint32_t max_doc, uint32_t* bad, cudaStream_t stream);

// result.cu
int result_merge(tagg_result* dst, const tagg_result* src);
// compact.cu: a lazily read result downloads its arrays on first use
int result_ensure_host(tagg_result* res);

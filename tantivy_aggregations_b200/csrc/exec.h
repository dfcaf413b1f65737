// exec.h — state of one tagg_execute call, shared by exec.cu / stream.cu / result.cu / comm.cu.
#pragma once
#include <string.h>

#include <cmath>

#include "host.h"

// host twins of the device codecs (dev.cuh)
inline double code_to_f64_h(uint64_t c) {
    uint64_t bits = (c >> 63) ? (c ^ 0x8000000000000000ull) : ~c;
    double d;
    memcpy(&d, &bits, 8);
    return d;
}
// host twin of dev.cuh hist_ord (IEEE double arithmetic is identical on both sides)
inline bool hist_ord_h(uint64_t code, double start, double interval, uint64_t* ord, uint32_t kind = TAGG_F64) {
    double k = kind == TAGG_F64 ? code_to_f64_h(code) : kind == TAGG_U64 ? (double)code : (double)(long long)(code ^ 0x8000000000000000ull);
    if (k != k) return false;
    volatile double n = k - start;
    if (n < 0.0) return false;
    volatile double q0 = n / interval;
    double q = std::floor(q0);
    if (!(q == q) || q <= 0.0) *ord = 0;
    else if (q >= 18446744073709551616.0) *ord = ~0ull;
    else *ord = (uint64_t)q;
    return true;
}
inline uint64_t f64_to_code_h(double v) {
    uint64_t bits;
    memcpy(&bits, &v, 8);
    return (bits >> 63) == 0 ? bits ^ 0x8000000000000000ull : ~bits;
}


struct ScopeLayout {
    int mode = SCOPE_DENSE;
    uint64_t capacity = 1;
    uint64_t dom_min = 0, dom_size = 1;
    size_t off_present = 0, off_keys = 0, off_parents = 0, off_state = 0, off_used = 0;
};
struct SlotLayout {
    size_t off_acc = 0, off_seen = 0;
    size_t off_edge = 0;  // exact f64 MIN / MAX (edge mode): 3 * capacity position cells, 0 = none
    uint64_t capacity = 1;
};

// compact.cu: the device-side image of the fruit
struct CompactState {
    uint8_t* d_img = nullptr;
    size_t d_bytes = 0;
    std::vector<uint64_t> cap_scope;
    std::vector<size_t> d_off_keys, d_off_parents, d_off_values, d_off_seen;
    std::vector<uint32_t*> d_rank;  // per scope: compact index of every raw cell (only where a child scope / nested percentile needs it)
    bool one_shot = false, launched = false, lazy = false;
};

struct ExecState {
    tagg_ctx* ctx = nullptr;
    const tagg_plan* plan = nullptr;
    const PlanMeta* meta = nullptr;
    cudaStream_t st = nullptr;
    CallRes* call = nullptr;
    bool collective = false;

    std::vector<const tagg_segment*> segs;
    std::vector<DevSegment> hsegs;
    DevSegment* d_segs = nullptr;
    std::vector<uint64_t> n_cand;    // candidates per segment (ids count or max_doc)
    std::vector<void*> temps;        // device allocations to release at the end
    uint32_t* d_bad_ids = nullptr;   // set by k_ids_to_bitset when a SORTED_IDS filter docset is malformed
    int mt_front_node = -1;          // TERMS node whose k_mterms launch measures its hot-key front (control words 2, 3)

    std::vector<ScopeLayout> scopes;
    std::vector<SlotLayout> slots;
    uint8_t* arena = nullptr;
    size_t arena_bytes = 0;
    size_t off_overflow = 0;
    size_t cls_begin[4] = {0, 0, 0, 0}, cls_end[4] = {0, 0, 0, 0};  // arena regions by reduction class
    DevPlan hplan;
    DevPlan* d_plan = nullptr;

    // percentile materialisation buffers (generic path)
    uint64_t* pct_codes[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* pct_buckets[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long* pct_count[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t pct_cap[4] = {0, 0, 0, 0};

    // percentiles on the streaming path (pct.cu): rank bins (count / min / max per equal-width code bin) between two
    // thresholds chosen from a sample, exact lists outside them
    struct RankState {
        bool active = false;
        uint64_t lo = 0, span = 0;   // binned codes: lo <= code < lo + span
        bool linear = false;         // bins equal-width in the VALUE: bin = (uint32)((f64(code) - f_lo) * f_scale), clamped
        double f_lo = 0, f_scale = 0;
        uint32_t shift = 0, mul = 0, n_bins = 0;  // bin = umulhi((code - lo) >> shift, mul), or (code - lo) >> shift if mul == 0
        uint8_t* d_block = nullptr;  // [tail counters 16 B][count u64 x n_bins][min][max][present u8 x n_bins]
        uint64_t *d_count = nullptr, *d_min = nullptr, *d_max = nullptr;
        uint8_t* d_present = nullptr;
        unsigned long long* d_tail_count = nullptr;  // [0] appended values, [1] those below lo
        uint64_t* d_tail = nullptr;
        uint64_t tail_cap = 0;
        PctSummary summary;          // filled by pct_rank_collect
        bool from_cache = false;     // thresholds reused from the plan (no sample pass)
        uint32_t node = 0;
        const uint8_t* h_block = nullptr;  // pinned copy of d_block, in flight before the pass's sync (pct_rank_prefetch)
        // With thresholds (and the list length) remembered from an earlier call, the exact lists are sorted and thinned on
        // the device right behind the pass, without waiting for their length on the host: the list is pre-filled with
        // all-ones codes and sorted at the predicted length (the fill sorts to the end), k_tail_pick reads the real
        // counters on the device.  One synchronisation per query; a list longer than predicted falls back to the host path.
        uint64_t tail_pred = 0;
        uint64_t* d_sorted = nullptr;
        void* d_cub = nullptr;
        size_t cub_bytes = 0;
        uint64_t* d_pick = nullptr;
        const uint64_t* h_pick = nullptr;
    } rank[4];
    bool no_rank = false;            // a rank-bin pass failed its precision check: redo on the exact path
    // f64 MIN / MAX follow the reference's PartialOrd fold (minmax.rs:97-106): per slot, 0 = the order of the codes is
    // exact, 1 = the column spans both zeros (exact unless the result is the ambiguous zero: checked after the read-out),
    // 2 = the column holds a NaN.  edge_exact: run the plan on the generic kernel with position tracking (generic.cu)
    std::vector<uint8_t> slot_edge;
    bool edge_exact = false;

    CompactState compact;
    uint64_t alg_bytes = 0;
    uint64_t direct_bytes = 0;  // host docset bytes the kernels read in place over PCIe
    uint32_t n_launches = 0;
    // SMs the persistent kernels leave free: a collective call's key-domain agreement (a one-CTA NCCL all-reduce) is in
    // flight while the pass runs; a pass that owns every SM leaves one of its own CTAs waiting behind that all-reduce, and
    // with tiles dealt statically the whole pass then ends that much later (2 GPUs: 1.82 ms instead of 1.54)
    uint32_t reserve_sms = 0;
    uint32_t launches_at_ev0 = 0;  // ev0 is re-recorded right before the first kernel of the pass (stream.cu)
    uint32_t path_used = 0;
    // chunked execute: host docsets are uploaded on a second stream, segments grouped into chunks; the kernels of
    // chunk c wait only for chunk c's uploads (chunk_ev[c]) while later chunks are still crossing PCIe
    struct PendingUpload { void* dst; const void* src; size_t bytes; uint32_t seg; uint32_t* scatter_words; uint64_t scatter_n; int slot; };
    // all host bitset docsets of a call live in ONE device block (zeroed once): segments of equal size then sit at a
    // constant stride, and if the caller's buffers do too (slices of one pinned buffer) a chunk crosses PCIe as a single
    // 2-D copy instead of one small copy per segment
    uint8_t* ds_block = nullptr;
    size_t ds_bytes = 0, ds_used = 0;
    std::vector<PendingUpload> uploads;  // host docsets: issued on call->st2 (the upload stream) after the allocations
    uint32_t n_chunks = 1;
    std::vector<uint32_t> chunk_begin;  // n_chunks + 1 segment indices
    std::vector<uint8_t> skip;  // per plan node: sub-tree already handled by a streaming launch
    int hash_shift = 0;  // growth applied to hash scopes after an overflow
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    ~ExecState();
    void free_temps();
    // scratch from the call slot's block cache (host.h CallRes::blocks); released (kept for the next call) by free_temps or
    // cache_release.  nullptr: out of memory.
    void* cache_alloc(size_t bytes);
    void cache_release(void* p);
    std::vector<void*> cached;
    // copy a small control block into pinned memory so its upload is asynchronous; falls back to `src`
    const void* pin(const void* src, size_t bytes);
};

int exec_run(const tagg_plan* plan, const tagg_segment_input* inputs, uint32_t n_inputs, int mode, int root,
             tagg_result** out);
struct tagg_pending;
int exec_begin(const tagg_plan* plan, const tagg_segment_input* inputs, uint32_t n_inputs, tagg_pending** out);
int exec_wait(tagg_pending* p, tagg_result** out);
// compact.cu: device accumulators -> compact image (kernels), its download, the result's directory
int compact_launch(ExecState& es);
int compact_download_begin(ExecState& es, tagg_result* res);
int compact_finish(ExecState& es, tagg_result* res);
void compact_release(ExecState& es);
// result.cu: percentile summaries of the exact path (after compact_finish)
int read_percentiles(ExecState& es, tagg_result* res);
// comm.cu: the key-domain agreement (one tiny min all-reduce per call, asynchronous: begin, then wait for the vector),
// the in-place merge of the accumulators of all ranks (dense scopes; root < 0: every rank receives the merged tables,
// else only `root`), and — for hashed tables / percentile summaries — the all-gather of the compact results folded on
// the host in rank order
int comm_agree_begin(ExecState& es, const std::vector<uint64_t>& dom);
int comm_agree_wait(ExecState& es, std::vector<uint64_t>& agreed);
int comm_merge_arena(ExecState& es, int root);
int comm_merge_results(ExecState& es, tagg_result* res);
// stream.cu: returns 1 if a streaming fast shape handled the plan, 0 if not applicable, <0 = -status
int stream_try(ExecState& es);
// mterms.cu: terms keyed by a multi-valued field / hashed key domain; returns members handled, <0 = -status
int mterms_try(ExecState& es);
// pct.cu: percentiles on the streaming path.  plan: 1 = rank-bin mode configured in es.rank[k], 0 = use the exact
// path, <0 = -status.  collect (after the pass, synchronises): 1 = summary ready, 0 = precision check failed.
int pct_rank_plan(ExecState& es, uint32_t node, int k);
int pct_rank_prefetch(ExecState& es);  // after the pass, before its sync: start the download of the bin tables
int pct_rank_collect(ExecState& es, int k);
void pct_rank_release(ExecState& es);

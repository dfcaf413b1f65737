// exec.cu — one `tagg_execute`: resolve columns, normalise docsets, size the bucket scopes, lay the
// accumulator arena out in HBM, launch the kernels, read the fruit back.
#include <string.h>

#include <algorithm>
#include <cmath>

#include <chrono>
#include <functional>
#include <stdio.h>
#include <stdlib.h>

#include "exec.h"

#define DENSE_MAX_BUCKETS (1ull << 24)
#define HASH_MIN_CAP (1ull << 10)
#define HASH_MAX_CAP (1ull << 30)

ExecState::~ExecState() {
    free_temps();
    if (call) {
        cudaStreamSynchronize(call->st2);  // host docsets are borrowed for the duration of the call only
        cudaStreamSynchronize(st);         // the pinned block may still feed an in-flight copy
        ctx->release_call(call);
    }
}
const void* ExecState::pin(const void* src, size_t bytes) {
    size_t at = (call->pinned_used + 63) & ~(size_t)63;
    if (!call->pinned || at + bytes > call->pinned_bytes) return src;
    if (src) memcpy(call->pinned + at, src, bytes);  // (src == nullptr: just a pinned scratch block)
    call->pinned_used = at + bytes;
    return call->pinned + at;
}
void* ExecState::cache_alloc(size_t bytes) {
    bytes = (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
    CallRes::DevBlock* best = nullptr;
    for (auto& b : call->blocks)
        if (!b.in_use && b.bytes >= bytes && b.bytes <= std::max<size_t>(2 * bytes, bytes + (1u << 20)) && (!best || b.bytes < best->bytes)) best = &b;
    if (!best) {
        if (call->blocks.size() >= 48) {  // keep the cache small: drop what is not in use
            std::vector<CallRes::DevBlock> keep;
            for (auto& b : call->blocks) { if (b.in_use) keep.push_back(b); else cudaFreeAsync(b.p, st); }
            call->blocks.swap(keep);
        }
        void* p = nullptr;
        if (cudaMallocAsync(&p, bytes, st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        call->blocks.push_back({p, bytes, false});
        best = &call->blocks.back();
    }
    best->in_use = true;
    cached.push_back(best->p);
    return best->p;
}
void ExecState::cache_release(void* p) {
    if (!p) return;
    for (auto& b : call->blocks)
        if (b.p == p) { b.in_use = false; break; }
    cached.erase(std::remove(cached.begin(), cached.end(), p), cached.end());
}
void ExecState::free_temps() {
    pct_rank_release(*this);
    compact_release(*this);
    for (void* p : temps) cudaFreeAsync(p, st);
    temps.clear();
    arena = nullptr;
    d_plan = nullptr;
    if (call) {
        for (void* p : cached)
            for (auto& b : call->blocks)
                if (b.p == p) { b.in_use = false; break; }
    }
    cached.clear();
    for (int i = 0; i < 4; i++) {
        if (pct_codes[i]) cudaFreeAsync(pct_codes[i], st);
        if (pct_buckets[i]) cudaFreeAsync(pct_buckets[i], st);
        if (pct_count[i]) cudaFreeAsync(pct_count[i], st);
        pct_codes[i] = nullptr; pct_buckets[i] = nullptr; pct_count[i] = nullptr;
    }
}

__global__ void k_fill_u64(uint64_t* __restrict__ dst, uint64_t n, uint64_t v) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = v;
}

static inline uint64_t pow2ceil(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}
static inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

template <typename T>
static int dev_alloc(ExecState& es, T** out, size_t bytes) {
    void* p = es.cache_alloc(bytes);
    if (!p) return tagg_fail(TAGG_ERR_OOM, "device scratch allocation failed (%zu bytes)", bytes);
    *out = (T*)p;
    return 0;
}

// A docset argument -> what the kernels test (dev.cuh DevDocset).
static int normalise_docset(ExecState& es, const tagg_segment* seg, uint32_t seg_index, const tagg_docset& in, bool is_main, int slot,
                            DevSegment& hs, int& next_col, DevDocset* out, uint64_t* n_cand) {
    DevDocset d;
    memset(&d, 0, sizeof(d));
    size_t need = ((size_t)seg->max_doc + 7) / 8;
    size_t words = (((size_t)seg->max_doc + TAGG_TILE_DOCS - 1) / TAGG_TILE_DOCS) * (TAGG_TILE_DOCS / 32) + 16;
    switch (in.kind) {
        case TAGG_DOCSET_ALL:
            d.kind = DS_ALL;
            if (n_cand) *n_cand = seg->max_doc;
            break;
        case TAGG_DOCSET_BITSET: {
            if (need && (!in.data || in.n < need))
                return tagg_fail(TAGG_ERR_BAD_ARG, "bitset docset needs %zu bytes for max_doc=%u, got %llu", need, seg->max_doc,
                                 (unsigned long long)in.n);
            {   // a page-locked, device-mapped host buffer (cudaHostAlloc / cudaHostRegister) is read IN PLACE: the streaming
                // kernel's TMA producer pulls each 256-byte tile over PCIe exactly once, overlapped with the column tiles —
                // no staging copy, no copy-engine latency.  Needs 16-byte alignment and room for whole 16-byte reads.
                static const bool no_direct = getenv("TAGG_NO_DIRECT") != nullptr;
                const size_t need16 = (need + 15) & ~(size_t)15;
                cudaPointerAttributes attr;
                if (!no_direct && need && in.n >= need16 && cudaPointerGetAttributes(&attr, in.data) == cudaSuccess &&
                    attr.type == cudaMemoryTypeHost && attr.devicePointer && ((uintptr_t)attr.devicePointer & 15) == 0) {
                    d.kind = DS_BITSET;
                    d.words = (const uint32_t*)attr.devicePointer;
                    d.n = need16;
                    if (n_cand) *n_cand = seg->max_doc;
                    es.alg_bytes += need;
                    es.direct_bytes += need;
                    break;
                }
                cudaGetLastError();
            }
            uint32_t* w = nullptr;
            const size_t take = (words * 4 + 255) & ~(size_t)255;
            if (es.ds_block && es.ds_used + take <= es.ds_bytes) {  // sub-allocated from the call's zeroed docset block
                w = (uint32_t*)(es.ds_block + es.ds_used);
                es.ds_used += take;
            } else {
                int rc = dev_alloc(es, &w, words * 4);
                if (rc) return rc;
                size_t tail = need / 4;
                CUDA_TRY(cudaMemsetAsync(w + tail, 0, (words - tail) * 4, es.st));
            }
            if (need) es.uploads.push_back({w, in.data, need, seg_index, nullptr, 0, slot});
            d.kind = DS_BITSET;
            d.words = w;
            if (n_cand) *n_cand = seg->max_doc;
            es.alg_bytes += need;
            break;
        }
        case TAGG_DOCSET_DEVICE_BITSET: {
            bool known = false;
            for (auto p : seg->cached_bitsets) known = known || p == in.data;
            if (!known) return tagg_fail(TAGG_ERR_BAD_ARG, "device bitset was not created by tagg_docset_cache on this segment");
            d.kind = DS_BITSET;
            d.words = (const uint32_t*)in.data;
            if (n_cand) *n_cand = seg->max_doc;
            es.alg_bytes += need;
            break;
        }
        case TAGG_DOCSET_SORTED_IDS: {
            if (in.n && !in.data) return tagg_fail(TAGG_ERR_BAD_ARG, "null doc-id list");
            uint32_t* ids = nullptr;
            int rc = dev_alloc(es, &ids, in.n * 4);
            if (rc) return rc;
            uint32_t* scatter = nullptr;
            if (!is_main) {  // filters are tested per doc: the ids are scattered into a bitset (K6) right after their upload
                rc = dev_alloc(es, &scatter, words * 4);
                if (rc) return rc;
                CUDA_TRY(cudaMemsetAsync(scatter, 0, words * 4, es.st));
                if (!es.d_bad_ids) {  // flag: an id >= max_doc / a list that is not strictly ascending
                    rc = dev_alloc(es, &es.d_bad_ids, 16);
                    if (rc) return rc;
                    CUDA_TRY(cudaMemsetAsync(es.d_bad_ids, 0, 16, es.st));
                }
            }
            if (in.n) es.uploads.push_back({ids, in.data, (size_t)in.n * 4, seg_index, scatter, in.n, slot});
            es.alg_bytes += in.n * 4;
            if (is_main) {
                d.kind = DS_IDS;
                d.ids = ids;
                d.n = in.n;
                if (n_cand) *n_cand = in.n;
            } else {
                d.kind = DS_BITSET;
                d.words = scatter;
            }
            break;
        }
        case TAGG_DOCSET_COLUMN_RANGE: {
            auto it = seg->cols.find(in.field_id);
            if (it == seg->cols.end())
                return tagg_fail(TAGG_ERR_NO_SUCH_COLUMN, "docset field %u is not a single-valued fast field of the segment", in.field_id);
            int slot = -1;
            for (int c = 0; c < next_col; c++)
                if (hs.cols[c].words == (const uint64_t*)it->second.dptr && hs.cols[c].num_bits == it->second.num_bits &&
                    hs.cols[c].min_value == it->second.min_value && it->second.dptr)
                    slot = c;
            if (slot < 0) {
                if (next_col >= TAGG_MAX_COLS) return tagg_fail(TAGG_ERR_BAD_PLAN, "too many device columns");
                slot = next_col++;
                hs.cols[slot] = it->second.dev();
                es.alg_bytes += it->second.payload_bytes + 16;
            }
            d.kind = DS_RANGE;
            d.col = slot;
            d.lo = in.lo;
            d.hi = in.hi;
            if (n_cand) *n_cand = seg->max_doc;
            break;
        }
        default: return tagg_fail(TAGG_ERR_BAD_ARG, "unknown docset kind %d", in.kind);
    }
    *out = d;
    return 0;
}

static int resolve_segments(ExecState& es, const tagg_segment_input* inputs, uint32_t n_inputs) {
    const PlanMeta& m = *es.meta;
    es.hsegs.resize(n_inputs);
    es.n_cand.assign(n_inputs, 0);
    // host docsets to upload?  then pipeline: chunks of segments, kernels of a chunk start as soon as its docsets landed
    {
        uint64_t h2d = 0;
        for (uint32_t i = 0; i < n_inputs; i++) {
            auto bytes = [](const tagg_docset& d) { return d.kind == TAGG_DOCSET_BITSET ? d.n : d.kind == TAGG_DOCSET_SORTED_IDS ? d.n * 4 : 0; };
            h2d += bytes(inputs[i].docset);
            for (uint32_t f = 0; f < m.n_filters && f < inputs[i].n_filters; f++) h2d += bytes(inputs[i].filters[f]);
        }
        // Bitsets that are slices of one host buffer (constant stride, equal segments) travel as ONE 2-D copy: on this
        // platform a PCIe copy has a ~60-90 us floor, so one big copy followed by one launch beats any chunking
        // (measured, C2 e2e: 1 chunk 0.63 ms, 2: 0.67, 4: 0.74, 8: 0.86)
        bool one_copy = n_inputs >= 2 && h2d >= (1u << 20);
        ptrdiff_t strides[1 + TAGG_MAX_FILTERS] = {0};
        for (uint32_t i = 0; one_copy && i + 1 < n_inputs; i++) {
            const tagg_segment_input &a = inputs[i], &b = inputs[i + 1];
            if (!a.segment || !b.segment || a.segment->max_doc != b.segment->max_doc) { one_copy = false; break; }
            auto strided = [&](const tagg_docset& x, const tagg_docset& y, ptrdiff_t* stride) {
                if (x.kind != y.kind) return false;
                if (x.kind == TAGG_DOCSET_SORTED_IDS) return false;
                if (x.kind != TAGG_DOCSET_BITSET) return true;
                const ptrdiff_t d = (const uint8_t*)y.data - (const uint8_t*)x.data;
                if (d < (ptrdiff_t)(((size_t)a.segment->max_doc + 7) / 8)) return false;
                if (*stride && *stride != d) return false;
                *stride = d;
                return true;
            };
            if (!strided(a.docset, b.docset, &strides[0])) one_copy = false;
            for (uint32_t f = 0; one_copy && f < m.n_filters && f < a.n_filters && f < b.n_filters; f++)
                if (!strided(a.filters[f], b.filters[f], &strides[1 + f])) one_copy = false;
        }
        if (one_copy) { es.n_chunks = 1; }
        else { static const char* ov = getenv("TAGG_CHUNKS"); const uint32_t maxc = ov ? (uint32_t)atoi(ov) : 4u; es.n_chunks = (h2d >= (1u << 20) && n_inputs >= 2) ? std::min<uint32_t>(std::min<uint32_t>(std::max<uint32_t>(maxc, 1u), 8u), n_inputs) : 1; }
        {   // one zeroed device block for every host bitset docset of the call
            size_t total = 0;
            auto words_of = [](const tagg_segment* sg) { return (((size_t)sg->max_doc + TAGG_TILE_DOCS - 1) / TAGG_TILE_DOCS) * (TAGG_TILE_DOCS / 32) + 16; };
            for (uint32_t i = 0; i < n_inputs; i++) {
                if (!inputs[i].segment) continue;
                const size_t take = (words_of(inputs[i].segment) * 4 + 255) & ~(size_t)255;
                if (inputs[i].docset.kind == TAGG_DOCSET_BITSET) total += take;
                for (uint32_t f = 0; f < m.n_filters && f < inputs[i].n_filters; f++)
                    if (inputs[i].filters[f].kind == TAGG_DOCSET_BITSET) total += take;
            }
            if (total) {
                void* blk = es.cache_alloc(total);
                if (!blk) return tagg_fail(TAGG_ERR_OOM, "docset staging allocation failed (%zu bytes)", total);
                CUDA_TRY(cudaMemsetAsync(blk, 0, total, es.st));
                es.ds_block = (uint8_t*)blk;
                es.ds_bytes = total;
                es.ds_used = 0;
            }
        }
        es.chunk_begin.assign(es.n_chunks + 1, 0);
        for (uint32_t c = 0; c <= es.n_chunks; c++) es.chunk_begin[c] = (uint32_t)((uint64_t)n_inputs * c / es.n_chunks);
    }
    for (uint32_t i = 0; i < n_inputs; i++) {
        const tagg_segment* seg = inputs[i].segment;
        if (!seg) return tagg_fail(TAGG_ERR_BAD_ARG, "input %u: null segment", i);
        if (seg->ctx != es.ctx) return tagg_fail(TAGG_ERR_BAD_ARG, "input %u: segment belongs to another context", i);
        if (inputs[i].n_filters < m.n_filters)
            return tagg_fail(TAGG_ERR_BAD_ARG, "input %u: plan has %u filter_agg nodes but %u filter docsets were given", i,
                             m.n_filters, inputs[i].n_filters);
        es.segs.push_back(seg);
        DevSegment& hs = es.hsegs[i];
        memset(&hs, 0, sizeof(hs));
        hs.max_doc = seg->max_doc;
        hs.has_deletes = seg->has_deletes ? 1 : 0;
        hs.deleted = seg->d_deleted;
        if (seg->has_deletes) es.alg_bytes += ((size_t)seg->max_doc + 7) / 8;
        int at = 0;
        for (auto& r : m.colrefs) {
            if (r.multi) {
                auto it = seg->mcols.find(r.field_id);
                if (it == seg->mcols.end())
                    return tagg_fail(TAGG_ERR_NO_SUCH_COLUMN, "field %u is not a multi-valued fast field of segment %u", r.field_id, i);
                hs.cols[at++] = it->second.first.dev();
                hs.cols[at++] = it->second.second.dev();
                es.alg_bytes += it->second.first.payload_bytes + it->second.second.payload_bytes + 32;
            } else {
                auto it = seg->cols.find(r.field_id);
                if (it == seg->cols.end())
                    return tagg_fail(TAGG_ERR_NO_SUCH_COLUMN, "field %u is not a single-valued fast field of segment %u", r.field_id, i);
                hs.cols[at++] = it->second.dev();
                es.alg_bytes += it->second.payload_bytes + 16;
            }
        }
        int next_col = at;
        int rc = normalise_docset(es, seg, i, inputs[i].docset, true, 0, hs, next_col, &hs.main, &es.n_cand[i]);
        if (rc) return rc;
        for (uint32_t f = 0; f < m.n_filters; f++) {
            rc = normalise_docset(es, seg, i, inputs[i].filters[f], false, 1 + (int)f, hs, next_col, &hs.filters[f], nullptr);
            if (rc) return rc;
        }
    }
    return 0;
}

// Host docsets cross PCIe on the upload stream, chunk by chunk; chunk_ev[c] fires when chunk c has landed.
static int issue_uploads(ExecState& es) {
    if (es.uploads.empty()) { es.n_chunks = 1; return 0; }
    cudaStream_t up = es.call->st2;
    CUDA_TRY(cudaEventRecord(es.call->join_ev, es.st));  // allocations / tail memsets issued so far
    CUDA_TRY(cudaStreamWaitEvent(up, es.call->join_ev, 0));
    size_t at = 0;
    for (uint32_t c = 0; c < es.n_chunks; c++) {
        const size_t c0 = at;
        while (at < es.uploads.size() && es.uploads[at].seg < es.chunk_begin[c + 1]) at++;
        std::vector<uint8_t> done(at - c0, 0);
        // same-slot bitsets of consecutive segments at constant source and destination strides: one 2-D copy
        for (size_t i = c0; i < at; i++) {
            if (done[i - c0] || es.uploads[i].scatter_words) continue;
            std::vector<size_t> run = {i};
            for (size_t j = i + 1; j < at; j++) {
                const auto &a = es.uploads[run.back()], &b = es.uploads[j];
                if (done[j - c0] || b.scatter_words || b.slot != a.slot || b.bytes != a.bytes || b.seg != a.seg + 1) continue;
                const ptrdiff_t ss = (const uint8_t*)b.src - (const uint8_t*)a.src, ds = (uint8_t*)b.dst - (uint8_t*)a.dst;
                if (ss < (ptrdiff_t)a.bytes || ds < (ptrdiff_t)a.bytes) break;
                if (run.size() >= 2) {
                    const auto& z = es.uploads[run[run.size() - 2]];
                    if (ss != (const uint8_t*)a.src - (const uint8_t*)z.src || ds != (uint8_t*)a.dst - (uint8_t*)z.dst) break;
                }
                run.push_back(j);
            }
            if (run.size() >= 2) {
                const auto &a = es.uploads[run[0]], &b = es.uploads[run[1]];
                CUDA_TRY(cudaMemcpy2DAsync(a.dst, (size_t)((uint8_t*)b.dst - (uint8_t*)a.dst), a.src, (size_t)((const uint8_t*)b.src - (const uint8_t*)a.src),
                                           a.bytes, run.size(), cudaMemcpyHostToDevice, up));
                for (size_t k : run) done[k - c0] = 1;
            }
        }
        for (size_t i = c0; i < at; i++) {
            if (done[i - c0]) continue;
            auto& u = es.uploads[i];
            CUDA_TRY(cudaMemcpyAsync(u.dst, u.src, u.bytes, cudaMemcpyHostToDevice, up));
            if (u.scatter_words) {
                CUDA_TRY(launch_ids_to_bitset((const uint32_t*)u.dst, u.scatter_n, u.scatter_words, es.segs[u.seg]->max_doc, es.d_bad_ids, up));
                es.ctx->launches++;
                es.n_launches++;
            }
        }
        CUDA_TRY(cudaEventRecord(es.call->chunk_ev[c], up));
    }
    return 0;
}

// Key domain of every bucket scope over the segments of this call (absolute codes / ordinals, so
// one table serves all segments — and, after comm_agree_domains, all ranks).
// dom[3*s+0] = smallest key (~0: none seen), [1] = ~(largest key), [2] = 1 if a dense table is possible.
static int scope_domains_local(ExecState& es, std::vector<uint64_t>& dom, std::vector<uint64_t>& bounds) {
    const PlanMeta& m = *es.meta;
    size_t ns = m.scope_node.size();
    dom.assign(ns * 3, 0);
    bounds.assign(ns, 0);
    for (size_t s = 1; s < ns; s++) {
        int node = m.scope_node[s];
        const tagg_node& nd = m.nodes[node];
        bool have = false;
        uint64_t lo = 0, hi = 0, bound = 0;
        for (size_t i = 0; i < es.segs.size(); i++) {
            const HostColumn* c = nd.multi ? &es.segs[i]->mcols.at(nd.field_id).second : &es.segs[i]->cols.at(nd.field_id);
            bound += nd.multi ? c->n_values : es.n_cand[i];
            if (c->n_values == 0) continue;
            uint64_t a = c->min_value, b = c->min_value + c->amplitude;
            lo = have ? std::min(lo, a) : a;
            hi = have ? std::max(hi, b) : b;
            have = true;
        }
        bounds[s] = bound;
        uint64_t dense_ok = 1, dmin = ~0ull, dmax = 0;
        if (have) {
            if (nd.op == TAGG_OP_TERMS) {
                dmin = lo;
                dmax = hi;
            } else {
                double start = nd.f0, interval = nd.f1;
                if (!(start == start) || !(interval > 0.0) || std::isinf(interval) || std::isinf(start)) {
                    dense_ok = 0;
                } else {
                    uint64_t olo = 0, ohi = 0;
                    bool vlo = hist_ord_h(lo, start, interval, &olo, nd.kind);
                    bool vhi = hist_ord_h(hi, start, interval, &ohi, nd.kind);
                    if (!vlo) olo = 0;  // the smallest valid k is >= start, whose ordinal is >= 0
                    if (!vhi) {
                        double khi = code_to_f64_h(hi);
                        ohi = (nd.kind == TAGG_F64 && khi != khi && (hi >> 63)) ? ~0ull : olo;  // +NaN codes lie above +inf; below start: nothing valid
                    }
                    if (ohi < olo) ohi = olo;
                    dmin = olo;
                    dmax = ohi;
                }
            }
        }
        dom[3 * s] = dmin;
        dom[3 * s + 1] = ~dmax;
        dom[3 * s + 2] = dense_ok;
    }
    return 0;
}

static int finalize_scopes(ExecState& es, const std::vector<uint64_t>& dom, const std::vector<uint64_t>& bounds) {
    const PlanMeta& m = *es.meta;
    size_t ns = m.scope_node.size();
    es.scopes.assign(ns, ScopeLayout());
    for (size_t s = 1; s < ns; s++) {
        ScopeLayout& L = es.scopes[s];
        const ScopeLayout& PL = es.scopes[m.scope_parent[s]];
        uint64_t dmin = dom[3 * s], dmax = ~dom[3 * s + 1];
        bool dense_ok = dom[3 * s + 2] != 0;
        if (dmin == ~0ull && dmax == 0) { dmin = 0; dmax = 0; }  // no values anywhere
        uint64_t dsize = dmax - dmin + 1;                         // 0 on full-range wrap
        if (dsize == 0) dense_ok = false;
        if (PL.mode != SCOPE_DENSE) dense_ok = false;
        if (dense_ok) {
            unsigned __int128 cap = (unsigned __int128)PL.capacity * dsize;
            if (cap > DENSE_MAX_BUCKETS) dense_ok = false;
        }
        if (dense_ok) {
            L.mode = SCOPE_DENSE;
            L.dom_min = dmin;
            L.dom_size = dsize;
            L.capacity = PL.capacity * dsize;
        } else {
            L.mode = SCOPE_HASH;
            uint64_t want = std::max<uint64_t>(bounds[s], 1) * 2;
            uint64_t cap = std::min<uint64_t>(std::max<uint64_t>(pow2ceil(want), HASH_MIN_CAP), 1ull << 22);
            cap <<= es.hash_shift;
            if (cap > HASH_MAX_CAP) return tagg_fail(TAGG_ERR_OOM, "bucket table would exceed %llu slots", (unsigned long long)HASH_MAX_CAP);
            L.capacity = cap;
        }
    }
    return 0;
}

// The arena is laid out by reduction class so that a multi-GPU merge is at most four all-reduces:
//   [control][hash scope tables][u8 max: bucket existence + Option flags][u64 sum: counts, integer sums]
//   [f64 sum][u64 max: min / max (MIN stored complemented)]
static int layout_arena(ExecState& es) {
    const PlanMeta& m = *es.meta;
    size_t off = 0;
    es.off_overflow = off;
    off += 16;
    for (size_t s = 0; s < es.scopes.size(); s++) {
        ScopeLayout& L = es.scopes[s];
        if (L.mode == SCOPE_HASH) {
            L.off_keys = off; off = align16(off + L.capacity * 8);
            L.off_parents = off; off = align16(off + L.capacity * 4);
            L.off_state = off; off = align16(off + L.capacity * 4);
            L.off_used = off; off += 16;
        }
    }
    es.slots.assign(m.slot_node.size(), SlotLayout());
    for (size_t k = 0; k < m.slot_node.size(); k++) es.slots[k].capacity = es.scopes[m.scope_of[m.slot_node[k]]].capacity;
    es.cls_begin[0] = off;
    for (size_t s = 0; s < es.scopes.size(); s++) {
        ScopeLayout& L = es.scopes[s];
        if (L.mode == SCOPE_DENSE) { L.off_present = off; off = align16(off + L.capacity); }
    }
    for (size_t k = 0; k < es.slots.size(); k++) { es.slots[k].off_seen = off; off = align16(off + es.slots[k].capacity); }
    es.cls_end[0] = off;
    for (int cls = 1; cls <= 3; cls++) {
        es.cls_begin[cls] = off;
        for (size_t k = 0; k < es.slots.size(); k++) {
            const tagg_node& nd = m.nodes[m.slot_node[k]];
            int c = (nd.op == TAGG_OP_COUNT || (nd.op == TAGG_OP_SUM && nd.kind != TAGG_F64)) ? 1 : (nd.op == TAGG_OP_SUM ? 2 : 3);
            if (c != cls) continue;
            es.slots[k].off_acc = off;
            off = align16(off + es.slots[k].capacity * 8);
        }
        es.cls_end[cls] = off;
    }
    // exact f64 MIN / MAX (edge mode): position cells behind the reduction classes (never merged cell by cell)
    for (size_t k = 0; k < es.slots.size(); k++) {
        es.slots[k].off_edge = 0;
        if (es.edge_exact && k < es.slot_edge.size() && es.slot_edge[k]) {
            es.slots[k].off_edge = off;
            off = align16(off + es.slots[k].capacity * 24);
        }
    }
    es.arena_bytes = off;
    void* p = es.cache_alloc(es.arena_bytes);
    if (!p) return tagg_fail(TAGG_ERR_OOM, "accumulator arena allocation failed (%zu bytes)", es.arena_bytes);
    es.arena = (uint8_t*)p;
    CUDA_TRY(cudaMemsetAsync(es.arena, 0, es.arena_bytes, es.st));
    // f64 sum cells start at -0.0: x + -0.0 == x for every x, so the first value "replaces" (sum.rs:95-102) and a sum
    // whose addends are all -0.0 stays -0.0 in whatever order the partial sums meet
    if (es.cls_end[2] > es.cls_begin[2]) {
        const uint64_t n = (es.cls_end[2] - es.cls_begin[2]) / 8;
        k_fill_u64<<<(unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)es.ctx->sm_count * 4), 256, 0, es.st>>>(
            (uint64_t*)(es.arena + es.cls_begin[2]), n, F64_NEG_ZERO_BITS);
        CUDA_TRY(cudaGetLastError());
        es.ctx->launches++;
        es.n_launches++;
    }
    return 0;
}

// Which f64 MIN / MAX leaves can see a NaN or both zeros?  The column headers are exact bounds of the stored codes.
static void edge_scan(ExecState& es, bool* any_straddle, bool* any_nan) {
    const PlanMeta& m = *es.meta;
    es.slot_edge.assign(m.slot_node.size(), 0);
    *any_straddle = *any_nan = false;
    for (size_t k = 0; k < m.slot_node.size(); k++) {
        const tagg_node& nd = m.nodes[m.slot_node[k]];
        if (nd.kind != TAGG_F64 || (nd.op != TAGG_OP_MIN && nd.op != TAGG_OP_MAX)) continue;
        uint8_t e = 0;
        bool neg = false, pos = false;  // over all segments: a value <= -0.0 / >= +0.0 exists
        for (const tagg_segment* sg : es.segs) {
            const HostColumn* c = nullptr;
            if (nd.multi) { auto it = sg->mcols.find(nd.field_id); if (it != sg->mcols.end()) c = &it->second.second; }
            else { auto it = sg->cols.find(nd.field_id); if (it != sg->cols.end()) c = &it->second; }
            if (!c || !c->n_values) continue;
            const uint64_t lo = c->min_value, hi = c->min_value + c->amplitude;
            if (lo < CODE_NEG_INF || hi > CODE_POS_INF) e = 2;
            neg = neg || lo <= CODE_NEG_ZERO;
            pos = pos || hi >= CODE_POS_ZERO;
        }
        if (!e && neg && pos) e = 1;
        es.slot_edge[k] = e;
        *any_straddle = *any_straddle || e == 1;
        *any_nan = *any_nan || e == 2;
    }
}

static int alloc_percentile_buffers(ExecState& es) {
    const PlanMeta& m = *es.meta;
    // percentile materialisation: capacity = every value that could be inserted
    for (size_t k = 0; k < m.pct_node.size(); k++) {
        if (es.rank[k].active) continue;  // handled by a streaming launch (pct.cu)
        const tagg_node& nd = m.nodes[m.pct_node[k]];
        uint64_t cap = 0;
        for (size_t i = 0; i < es.segs.size(); i++)
            cap += nd.multi ? es.segs[i]->mcols.at(nd.field_id).second.n_values : es.n_cand[i];
        es.pct_cap[k] = cap;
        void* p = nullptr;
        CUDA_TRY(cudaMallocAsync(&p, cap * 8 + 16, es.st)); es.pct_codes[k] = (uint64_t*)p;
        CUDA_TRY(cudaMallocAsync(&p, cap * 4 + 16, es.st)); es.pct_buckets[k] = (uint32_t*)p;
        CUDA_TRY(cudaMallocAsync(&p, 16, es.st)); es.pct_count[k] = (unsigned long long*)p;
        CUDA_TRY(cudaMemsetAsync(es.pct_count[k], 0, 16, es.st));
    }
    return 0;
}

// every member of the plan's top tuple (below the leading filters) was handled by a fast-path launch
static bool plan_fully_covered(const ExecState& es) {
    const PlanMeta& m = *es.meta;
    uint32_t node = 0;
    while (node < m.nodes.size() && (m.nodes[node].op == TAGG_OP_FILTER || m.nodes[node].op == TAGG_OP_POST_FILTER)) node++;
    if (node >= m.nodes.size()) return false;
    if (m.nodes[node].op != TAGG_OP_TUPLE) return es.skip[node] != 0;
    for (uint32_t c = node + 1; c < m.end[node]; c = m.end[c])
        if (!es.skip[c]) return false;
    return true;
}

static int build_dev_plan(ExecState& es) {
    const PlanMeta& m = *es.meta;
    DevPlan& P = es.hplan;
    memset(&P, 0, sizeof(P));
    P.n_nodes = (uint32_t)m.nodes.size();
    P.n_scopes = (uint32_t)es.scopes.size();
    P.n_slots = (uint32_t)es.slots.size();
    for (int i = 0; i < TAGG_MAX_NODES; i++) P.slot_root_index[i] = -1;
    for (uint32_t i = 0; i < P.n_nodes; i++) {
        const tagg_node& nd = m.nodes[i];
        DevNode& d = P.nodes[i];
        d.op = nd.op; d.kind = nd.kind; d.multi = nd.multi; d.pred = nd.pred;
        d.col = (uint16_t)(m.col_slot[i] < 0 ? 0 : m.col_slot[i]);
        d.end = m.end[i];
        d.scope = (uint16_t)m.scope_of[i];
        d.own_scope = (uint16_t)(m.own_scope[i] < 0 ? 0 : m.own_scope[i]);
        d.slot = (uint16_t)(m.slot_of[i] < 0 ? 0 : m.slot_of[i]);
        d.aux = (uint16_t)(nd.op == TAGG_OP_PERCENTILES ? m.pct_of[i] : nd.aux);
        d.skip = i < es.skip.size() ? es.skip[i] : 0;
        d.lut = (nd.op == TAGG_OP_POST_FILTER && nd.pred == TAGG_PRED_LUT) ? es.plan->d_blobs[nd.aux] : nullptr;
        d.f0 = nd.f0; d.f1 = nd.f1; d.u0 = nd.u0; d.u1 = nd.u1;
    }
    for (size_t s = 0; s < es.scopes.size(); s++) {
        const ScopeLayout& L = es.scopes[s];
        DevScope& d = P.scopes[s];
        d.mode = L.mode;
        d.parent = m.scope_parent[s];
        d.capacity = L.capacity;
        d.dom_min = L.dom_min;
        d.dom_size = L.dom_size;
        if (L.mode == SCOPE_DENSE) {
            d.present = es.arena + L.off_present;
        } else {
            d.keys = (uint64_t*)(es.arena + L.off_keys);
            d.parents = (uint32_t*)(es.arena + L.off_parents);
            d.state = (uint32_t*)(es.arena + L.off_state);
            d.used = (unsigned long long*)(es.arena + L.off_used);
        }
    }
    uint32_t nroot = 0;
    for (size_t k = 0; k < es.slots.size(); k++) {
        P.slots[k].acc = (uint64_t*)(es.arena + es.slots[k].off_acc);
        P.slots[k].seen = es.arena + es.slots[k].off_seen;
        P.slots[k].edge = es.slots[k].off_edge ? (uint64_t*)(es.arena + es.slots[k].off_edge) : nullptr;
        P.slots[k].edge_cap = es.slots[k].capacity;
        int node = m.slot_node[k];
        if (m.scope_of[node] == 0 && nroot < TAGG_MAX_ROOT_SLOTS && !P.slots[k].edge) {
            P.slot_root_index[k] = (int16_t)nroot;
            P.root_slot_nodes[nroot++] = (uint16_t)node;
        }
    }
    P.n_root_slots = nroot;
    P.overflow = (uint32_t*)(es.arena + es.off_overflow);
    for (size_t k = 0; k < m.pct_node.size(); k++) {
        P.pct_codes[k] = es.pct_codes[k];
        P.pct_buckets[k] = es.pct_buckets[k];
        P.pct_count[k] = es.pct_count[k];
        P.pct_cap[k] = es.pct_cap[k];
    }
    void* p = es.cache_alloc(sizeof(DevPlan));
    if (!p) return tagg_fail(TAGG_ERR_OOM, "plan allocation failed");
    es.d_plan = (DevPlan*)p;
    CUDA_TRY(cudaMemcpyAsync(es.d_plan, es.pin(&P, sizeof(DevPlan)), sizeof(DevPlan), cudaMemcpyHostToDevice, es.st));
    return 0;
}

// One call of the hot path.  mode: 0 = this GPU only (tagg_execute), 1 = collective, every rank receives the merged fruit
// (tagg_execute_collective), 2 = collective, merged on `root` only (tagg_execute_reduce).
//
// The call is split at its ONE synchronisation point: begin() resolves the inputs, lays the arena out, launches the pass,
// the cross-GPU merge, the compaction and the download, and returns without waiting; finish() waits, checks the flags,
// redoes the pass in the rare cases that need it (hash table growth, percentile bins that failed their precision check,
// an ambiguous signed zero, a key-domain agreement that moved) and hands the fruit out.  tagg_execute = begin + finish;
// tagg_execute_begin / tagg_pending_wait let a host keep two queries in flight (each on its own stream and pinned block),
// so that the host-side preparation of query i+1 overlaps the kernels of query i.
struct ExecCall {
    ExecState es;
    const tagg_plan* plan = nullptr;
    tagg_ctx* ctx = nullptr;
    uint32_t n_inputs = 0;
    int mode = 0, root = -1;
    bool collective = false;
    bool edge_straddle = false, edge_nan = false;
    std::vector<uint64_t> agreed, dom_used;
    bool agree_pending = false;
    float ms_total = 0;
    tagg_result* res = nullptr;
    bool arena_merge = false, i_read = true, ok = false;
    uint32_t* flags = nullptr;
    uint32_t flags_local[4] = {0, 0, 0, 0};
    int attempt = 0;
    std::chrono::steady_clock::time_point t_begin;

    void lap(const char* what) {
        static const bool trace = getenv("TAGG_TRACE") != nullptr;
        if (!trace) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[tagg] %-18s %8.1f us\n", what, std::chrono::duration<double, std::micro>(now - t_begin).count());
    }
    void take_result() {
        {
            std::lock_guard<std::mutex> g(ctx->mu);
            if (!ctx->result_pool.empty()) { res = ctx->result_pool.back(); ctx->result_pool.pop_back(); }
        }
        if (!res) res = new tagg_result();
        res->ctx = ctx;
        res->meta = plan->meta;
        res->has_img = false;
        res->lazy = false;
        res->merged_elsewhere = 0;
        res->d_stream = es.st;
        res->pcts.clear();
    }
    void drop_result() {
        if (!res) return;
        res->release_device();
        res->meta.reset();
        res->pcts.clear();
        std::lock_guard<std::mutex> g(ctx->mu);
        if (ctx->result_pool.size() < 4) ctx->result_pool.push_back(res); else delete res;
        res = nullptr;
    }
    ~ExecCall() { if (!ok) drop_result(); }

    int begin(const tagg_plan* plan_, const tagg_segment_input* inputs, uint32_t n_inputs_, int mode_, int root_);
    int issue();                // one attempt, up to (not including) the synchronisation
    int complete(bool* redo);   // the synchronisation and everything behind it
    int finish(tagg_result** out);
};

int ExecCall::begin(const tagg_plan* plan_, const tagg_segment_input* inputs, uint32_t n_inputs_, int mode_, int root_) {
    plan = plan_; n_inputs = n_inputs_; mode = mode_; root = root_;
    ctx = plan->ctx;
    t_begin = std::chrono::steady_clock::now();
    CUDA_TRY(cudaSetDevice(ctx->device));
    collective = mode != 0;
    if (collective && !ctx->nccl) return tagg_fail(TAGG_ERR_NCCL, "collective execution needs tagg_comm_init first");
    if (mode == 2 && (root < 0 || root >= ctx->n_ranks)) return tagg_fail(TAGG_ERR_BAD_ARG, "root rank %d out of range", root);
    es.ctx = ctx;
    es.plan = plan;
    es.meta = plan->meta.get();
    es.collective = collective;
    es.call = ctx->acquire_call();
    es.st = es.call->st;
    es.ev0 = es.call->ev0;
    es.ev1 = es.call->ev1;
    lap("acquire");
    int rc = resolve_segments(es, inputs, n_inputs);
    if (rc) return rc;
    lap("resolve_segments");
    if (n_inputs) {
        rc = dev_alloc(es, &es.d_segs, sizeof(DevSegment) * n_inputs);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(es.d_segs, es.pin(es.hsegs.data(), sizeof(DevSegment) * n_inputs), sizeof(DevSegment) * n_inputs, cudaMemcpyHostToDevice, es.st));
    }
    rc = issue_uploads(es);
    if (rc) return rc;
    // f64 MIN / MAX: the order of the codes is the reference's PartialOrd fold except around NaN and the two zeros
    edge_scan(es, &edge_straddle, &edge_nan);
    es.edge_exact = edge_nan;  // a NaN in a column: exact path right away; zeros: only if the result turns out ambiguous
    attempt = 0;
    return issue();
}

// Collective: every rank lays its tables out over the SAME key domains, agreed with one tiny min all-reduce per call
// (comm.cu).  The agreed vector of the previous call on the same (plan, segment set) is used OPTIMISTICALLY: the pass
// starts at once on it while the all-reduce is in flight, and is redone in the rare case the agreement moved (some
// rank's segments changed).  Every rank issues exactly one agreement per call, so the NCCL call sequences always match.
int ExecCall::issue() {
    int rc = 0;
    for (;; attempt++) {
        std::vector<uint64_t> dom, bounds;
        rc = scope_domains_local(es, dom, bounds);
        if (rc) return rc;
        if (collective) {
            // one more agreed word: does ANY rank need the exact f64 MIN / MAX path?  (min-reduced: 0 = yes)
            dom.push_back((edge_straddle || edge_nan) ? 0ull : 1ull);
            if (attempt == 0) {
                rc = comm_agree_begin(es, dom);
                if (rc) return rc;
                agree_pending = true;
                std::vector<const void*> key(es.segs.begin(), es.segs.end());
                bool hit = false;
                {
                    std::lock_guard<std::mutex> g(plan->mu);
                    if (plan->dom_key == key && plan->dom_local == dom && plan->dom_agreed.size() == dom.size()) { dom_used = plan->dom_agreed; hit = true; }
                }
                if (!hit) {
                    rc = comm_agree_wait(es, agreed);
                    if (rc) return rc;
                    agree_pending = false;
                    dom_used = agreed;
                    std::lock_guard<std::mutex> g(plan->mu);
                    plan->dom_key = key; plan->dom_local = dom; plan->dom_agreed = agreed;
                }
            } else {
                dom_used = agreed;  // (a redo never issues a second agreement: the other ranks would not)
            }
            dom = dom_used;
            if (dom.back() == 0) es.edge_exact = true;
            dom.pop_back();
        }
        rc = finalize_scopes(es, dom, bounds);
        if (rc) return rc;
        rc = layout_arena(es);
        if (rc) return rc;
        lap("layout");

        {
            cudaEvent_t prev = nullptr;
            {
                std::lock_guard<std::mutex> g(ctx->mu);
                prev = ctx->last_pass_done;
            }
            if (prev && prev != es.call->chain_ev) CUDA_TRY(cudaStreamWaitEvent(es.st, prev, 0));
        }
        CUDA_TRY(cudaEventRecord(es.ev0, es.st));
        es.launches_at_ev0 = es.n_launches;
        es.reserve_sms = agree_pending ? 2u : 0u;
        es.skip.assign(es.meta->nodes.size(), 0);
        const bool fast_ok = ctx->path != 1 && !es.edge_exact;
        if (es.edge_exact && ctx->path == 2)
            return tagg_fail(TAGG_ERR_UNSUPPORTED, "f64 min / max over NaN or signed zeros runs on the exact (generic) path; path is forced to stream");
        int handled = 0;
        if (fast_ok) {
            handled = stream_try(es);
            if (handled < 0) return -handled;
        }
        // a rank-bin pass that was planned but whose launch did not happen (its member is not flagged) must not shadow
        // the exact path
        for (size_t k = 0; k < es.meta->pct_node.size() && k < 4; k++)
            if (es.rank[k].active && !es.skip[es.meta->pct_node[k]]) {
                if (es.rank[k].d_block) cudaFreeAsync(es.rank[k].d_block, es.st);
                if (es.rank[k].d_tail) cudaFreeAsync(es.rank[k].d_tail, es.st);
                if (es.rank[k].d_sorted) cudaFreeAsync(es.rank[k].d_sorted, es.st);
                if (es.rank[k].d_cub) cudaFreeAsync(es.rank[k].d_cub, es.st);
                if (es.rank[k].d_pick) cudaFreeAsync(es.rank[k].d_pick, es.st);
                es.rank[k] = ExecState::RankState();
            }
        int mt = 0;
        if (handled != 1 && fast_ok) {  // K5: terms keyed by multi-valued / hashed fields
            mt = mterms_try(es);
            if (mt < 0) return -mt;
            if (mt > 0 && plan_fully_covered(es)) handled = 1;
            else if (mt > 0) handled = 2;
        }
        if (handled != 1) {
            if (ctx->path == 2) return tagg_fail(TAGG_ERR_UNSUPPORTED, "the plan has no streaming fast shape (path forced to stream)");
            es.path_used = mt > 0 ? 5 : handled == 2 ? 3 : 1;
            rc = alloc_percentile_buffers(es);
            if (rc) return rc;
            rc = build_dev_plan(es);
            if (rc) return rc;
            if (!es.uploads.empty())
                for (uint32_t c = 0; c < es.n_chunks; c++) CUDA_TRY(cudaStreamWaitEvent(es.st, es.call->chunk_ev[c], 0));
            for (uint32_t i = 0; i < n_inputs; i++) {
                CUDA_TRY(launch_generic(es.d_plan, es.d_segs + i, es.n_cand[i], i, ctx->sm_count, es.st));
                if (es.n_cand[i]) { ctx->launches++; es.n_launches++; }
            }
            for (size_t k = 0; k < es.slots.size(); k++) {  // settle the exact f64 MIN / MAX cells (generic.cu k_edge_fixup)
                if (!es.slots[k].off_edge || !n_inputs) continue;
                const int node = es.meta->slot_node[k];
                const tagg_node& nd = es.meta->nodes[node];
                CUDA_TRY(launch_edge_fixup(es.d_segs, es.meta->col_slot[node] + (nd.multi ? 1 : 0), nd.op == TAGG_OP_MIN,
                                           (uint64_t*)(es.arena + es.slots[k].off_acc), es.arena + es.slots[k].off_seen,
                                           (const uint64_t*)(es.arena + es.slots[k].off_edge), es.slots[k].capacity, ctx->sm_count, es.st));
                ctx->launches++;
                es.n_launches++;
            }
        } else {
            es.path_used = mt > 0 ? 4 : 2;
        }
        CUDA_TRY(cudaEventRecord(es.ev1, es.st));
        lap("launched");

        // the optimistic layout: has the agreement confirmed it?
        if (agree_pending) {
            rc = comm_agree_wait(es, agreed);
            if (rc) return rc;
            agree_pending = false;
            if (agreed != dom_used) {  // some rank's segments changed: every rank sees the same new vector and redoes the pass
                {
                    std::lock_guard<std::mutex> g(plan->mu);
                    plan->dom_agreed = agreed;
                }
                CUDA_TRY(cudaStreamSynchronize(es.st));
                pct_rank_release(es);
                es.cache_release(es.arena); es.arena = nullptr;
                es.cache_release(es.d_plan);
                es.d_plan = nullptr;
                continue;
            }
        }
        // dense tables merge cell by cell on the device, BEFORE the read-out (no host round trip in between: a table that
        // can be merged this way cannot overflow); hashed tables, percentile summaries and exact f64 MIN / MAX cells (NaN /
        // zero order is per rank) as compact results, after it
        arena_merge = collective && es.meta->pct_node.empty() && !es.edge_exact;
        for (auto& L : es.scopes) arena_merge = arena_merge && L.mode == SCOPE_DENSE;
        i_read = !(arena_merge && mode == 2 && ctx->rank != root);
        if (arena_merge) {
            rc = comm_merge_arena(es, mode == 2 ? root : -1);
            if (rc) return rc;
        }
        rc = pct_rank_prefetch(es);
        if (rc) return rc;
        flags = (uint32_t*)const_cast<void*>(es.pin(nullptr, 16));  // [overflow][bad ids]
        if (!flags) flags = flags_local;
        flags[0] = flags[1] = 0;
        take_result();
        if (i_read) {
            rc = compact_launch(es);
            if (rc) return rc;
        }
        CUDA_TRY(cudaEventRecord(es.call->chain_ev, es.st));
        {
            std::lock_guard<std::mutex> g(ctx->mu);
            ctx->last_pass_done = es.call->chain_ev;
        }
        if (i_read) {
            rc = compact_download_begin(es, res);
            if (rc) return rc;
        }
        CUDA_TRY(cudaMemcpyAsync(&flags[0], es.arena + es.off_overflow, 4, cudaMemcpyDeviceToHost, es.st));
        if (es.d_bad_ids) CUDA_TRY(cudaMemcpyAsync(&flags[1], es.d_bad_ids, 4, cudaMemcpyDeviceToHost, es.st));
        flags[2] = flags[3] = 0;
        if (es.mt_front_node >= 0 && flags != flags_local) CUDA_TRY(cudaMemcpyAsync(&flags[2], es.arena + es.off_overflow + 8, 8, cudaMemcpyDeviceToHost, es.st));
        return 0;
    }
}

int ExecCall::complete(bool* redo_out) {
    int rc = 0;
    *redo_out = false;
    CUDA_TRY(cudaStreamSynchronize(es.st));
    lap("synced");
    if (flags[1]) return tagg_fail(TAGG_ERR_BAD_ARG, "a sorted-id filter docset holds ids that are not strictly ascending or >= max_doc of its segment");
    float ms = 0;
    cudaEventElapsedTime(&ms, es.ev0, es.ev1);
    ms_total += ms;
    uint32_t overflow = flags[0];
    bool redo = overflow != 0;
    if (es.mt_front_node >= 0 && flags[2]) {  // the hot-key front's measured hit rate (mterms.cu): remembered with the plan
        std::lock_guard<std::mutex> g(es.plan->mu);
        es.plan->mt_front_hint[es.mt_front_node] = (uint64_t)flags[3] * 64u >= flags[2] ? 1 : 2;
    }
    if (!overflow) {
        for (int k = 0; k < 4 && !redo; k++)
            if (es.rank[k].active) {
                const int pr = pct_rank_collect(es, k);
                if (pr < 0) return -pr;
                if (pr == 0) { redo = true; overflow = 3; }
            }
    }
    if (!redo) {
        if (i_read) {
            rc = compact_finish(es, res);
            if (rc) return rc;
            rc = read_percentiles(es, res);
            if (rc) return rc;
        } else {
            res->merged_elsewhere = 1;
            res->n_scope.assign(es.scopes.size(), 0);
            res->off_keys.assign(es.scopes.size(), 0); res->off_parents.assign(es.scopes.size(), 0);
            res->off_values.assign(es.slots.size(), 0); res->off_seen.assign(es.slots.size(), 0);
            if (!res->img) { res->img = (uint8_t*)malloc(64); res->img_cap = 64; res->img_pinned = false; }
            if (!res->img) return tagg_fail(TAGG_ERR_OOM, "out of memory");
            res->has_img = true;
            res->pcts.resize(es.meta->pct_node.size());
        }
        // a column that spans both zeros: the order of the codes picked -0.0 as the minimum (+0.0 as the maximum); if
        // the other zero was collected too the reference keeps whichever came FIRST (minmax.rs:99-102) — exact path
        if (!es.edge_exact && edge_straddle && !collective) {
            rc = result_ensure_host(res);
            if (rc) return rc;
            bool ambiguous = false;
            for (size_t k = 0; k < es.slot_edge.size() && !ambiguous; k++) {
                if (es.slot_edge[k] != 1) continue;
                const uint64_t amb = es.meta->nodes[es.meta->slot_node[k]].op == TAGG_OP_MIN ? F64_NEG_ZERO_BITS : 0ull;
                const uint64_t n = res->slot_len(k);
                const uint64_t* V = res->slot_values(k);
                const uint8_t* Sn = res->slot_seen(k);
                for (size_t i = 0; i < n && !ambiguous; i++) ambiguous = Sn[i] && V[i] == amb;
            }
            if (ambiguous) {
                if (ctx->path == 2) return tagg_fail(TAGG_ERR_UNSUPPORTED, "f64 min / max over both signed zeros runs on the exact (generic) path; path is forced to stream");
                es.edge_exact = true;
                redo = true;
            }
        }
        if (!redo) return 0;
    }
    drop_result();
    if (overflow == 3) {  // the rank bins could not resolve this distribution
        bool cached = false;
        for (int k = 0; k < 4; k++) cached = cached || (es.rank[k].active && es.rank[k].from_cache);
        if (cached) {  // ... with thresholds remembered from another docset: sample this one afresh
            std::lock_guard<std::mutex> g(plan->mu);
            for (auto& pc : plan->pct_cache) pc.valid = false;
        } else {
            es.no_rank = true;  // exact path
        }
    }
    if (overflow == 2) return tagg_fail(TAGG_ERR_CUDA, "percentile buffer overflow (internal sizing error)");
    if (overflow == 4) return tagg_fail(TAGG_ERR_BAD_ARG, "a column holds values outside the range its header declares (min_value / num_bits)");
    if (overflow == 5) return tagg_fail(TAGG_ERR_BAD_ARG, "a sorted-id docset holds ids that are not strictly ascending or >= max_doc of its segment");
    if (attempt >= 7) return tagg_fail(TAGG_ERR_OOM, "bucket table kept overflowing");
    // a hash scope ran out of room: grow 4x and redo the pass from clean accumulators
    if (overflow == 1) es.hash_shift += 2;
    pct_rank_release(es);
    compact_release(es);
    es.cache_release(es.arena); es.arena = nullptr;
    es.cache_release(es.d_plan);
    es.d_plan = nullptr;
    for (int k = 0; k < 4; k++) {
        if (es.pct_codes[k]) cudaFreeAsync(es.pct_codes[k], es.st);
        if (es.pct_buckets[k]) cudaFreeAsync(es.pct_buckets[k], es.st);
        if (es.pct_count[k]) cudaFreeAsync(es.pct_count[k], es.st);
        es.pct_codes[k] = nullptr; es.pct_buckets[k] = nullptr; es.pct_count[k] = nullptr;
    }
    attempt++;
    *redo_out = true;
    return 0;
}

int ExecCall::finish(tagg_result** out) {
    int rc = 0;
    for (;;) {
        bool redo = false;
        rc = complete(&redo);
        if (rc) return rc;
        if (!redo) break;
        rc = issue();
        if (rc) return rc;
    }
    if (collective && !arena_merge) {
        rc = comm_merge_results(es, res);
        if (rc) return rc;
    }
    lap("read_result");
    res->kernel_ms = ms_total;
    res->alg_bytes = es.alg_bytes;
    res->n_launches = es.n_launches;
    res->path_used = es.path_used;
    *out = res;
    ok = true;
    return 0;
}

int exec_run(const tagg_plan* plan, const tagg_segment_input* inputs, uint32_t n_inputs, int mode, int root, tagg_result** out) {
    if (!plan || !out || (n_inputs && !inputs)) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_execute: null argument");
    ExecCall call;
    int rc = call.begin(plan, inputs, n_inputs, mode, root);
    if (rc) return rc;
    return call.finish(out);
}

// ---- two queries in flight: begin now, wait later -----------------------------------------------------------------
struct tagg_pending { ExecCall call; };

int exec_begin(const tagg_plan* plan, const tagg_segment_input* inputs, uint32_t n_inputs, tagg_pending** out) {
    if (!plan || !out || (n_inputs && !inputs)) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_execute_begin: null argument");
    auto* p = new tagg_pending();
    int rc = p->call.begin(plan, inputs, n_inputs, 0, -1);
    if (rc) { delete p; return rc; }
    *out = p;
    return 0;
}
int exec_wait(tagg_pending* p, tagg_result** out) {
    if (!p || !out) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_pending_wait: null argument");
    int rc = p->call.finish(out);
    delete p;
    return rc;
}

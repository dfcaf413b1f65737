// result.cu — device accumulators -> compact host fruit, the result readers of the ABI, and
// `PreparedAgg::merge` (count.rs:39-41, sum.rs:59-70, minmax.rs:59-72, terms.rs:85-92,
// histogram.rs:90-97, percentile.rs:58-62) on compact results.
#include <string.h>

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <unordered_map>
#include <chrono>
#include <stdio.h>
#include <stdlib.h>

#include "exec.h"

__global__ void k_gather_ranks(const uint64_t* __restrict__ sorted, const uint64_t* __restrict__ ranks, uint64_t n,
                               uint64_t* __restrict__ dst) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = sorted[ranks[i] - 1];
}

static inline uint64_t code_to_bits_h(int kind, uint64_t c) {
    if (kind == TAGG_U64) return c;
    if (kind == TAGG_F64) return (c >> 63) ? (c ^ 0x8000000000000000ull) : ~c;
    return c ^ 0x8000000000000000ull;
}

// Ranks kept from an exactly sorted multiset of n values: every rank up to 4096, then a geometric
// schedule with ratio 1 + eps/4 (eps = 0.01, percentile.rs:174) so that any target rank k has a
// stored rank within eps*k/8 — always inside CKMS's own +-eps*k band.
static void rank_schedule(uint64_t n, std::vector<uint64_t>& ranks, uint64_t dense_upto = 4096, double ratio = 0.0025) {
    ranks.clear();
    uint64_t dense = std::min<uint64_t>(n, dense_upto);
    for (uint64_t r = 1; r <= dense; r++) ranks.push_back(r);
    uint64_t r = dense;
    while (r < n) {
        uint64_t nx = r + std::max<uint64_t>(1, (uint64_t)((double)r * ratio));
        if (nx > n) nx = n;
        ranks.push_back(nx);
        r = nx;
    }
}

// percentiles under a bucket aggregation: the materialised (code, bucket) pairs are sorted by code, then stably by
// bucket; every run of one bucket is an exactly sorted multiset of which a rank schedule is kept (every rank up to 256,
// then geometric with ratio 1 + eps/4 — the stored neighbour of any target rank is within eps/8 of it)
static int read_nested_percentiles(ExecState& es, tagg_result* res, size_t k, uint64_t n, const std::vector<uint32_t>& rank_of_raw) {
    if (n > 0x7fffffffull) return tagg_fail(TAGG_ERR_UNSUPPORTED, "more than 2^31-1 percentile values in one call (%llu): split the call by segment and merge", (unsigned long long)n);
    uint64_t* d_codes_alt = nullptr;
    uint32_t* d_bkt_alt = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&d_codes_alt, n * 8, es.st)); es.temps.push_back(d_codes_alt);
    CUDA_TRY(cudaMallocAsync((void**)&d_bkt_alt, n * 4, es.st)); es.temps.push_back(d_bkt_alt);
    size_t t1 = 0, t2 = 0, t3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 64, es.st);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)n, 0, 32, es.st);
    uint32_t *d_unique = nullptr, *d_counts = nullptr, *d_nruns = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&d_unique, n * 4, es.st)); es.temps.push_back(d_unique);
    CUDA_TRY(cudaMallocAsync((void**)&d_counts, n * 4, es.st)); es.temps.push_back(d_counts);
    CUDA_TRY(cudaMallocAsync((void**)&d_nruns, 16, es.st)); es.temps.push_back(d_nruns);
    cub::DeviceRunLengthEncode::Encode(nullptr, t3, (const uint32_t*)nullptr, d_unique, d_counts, d_nruns, (int)n, es.st);
    void* d_tmp = nullptr;
    const size_t tb = std::max(t1, std::max(t2, t3)) + 16;
    CUDA_TRY(cudaMallocAsync(&d_tmp, tb, es.st)); es.temps.push_back(d_tmp);
    size_t tt = tb;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tt, (const uint64_t*)es.pct_codes[k], d_codes_alt, (const uint32_t*)es.pct_buckets[k], d_bkt_alt, (int)n, 0, 64, es.st));
    tt = tb;  // stable: codes stay ascending inside a bucket
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tt, (const uint32_t*)d_bkt_alt, es.pct_buckets[k], (const uint64_t*)d_codes_alt, es.pct_codes[k], (int)n, 0, 32, es.st));
    tt = tb;
    CUDA_TRY(cub::DeviceRunLengthEncode::Encode(d_tmp, tt, (const uint32_t*)es.pct_buckets[k], d_unique, d_counts, d_nruns, (int)n, es.st));
    uint32_t nruns = 0;
    CUDA_TRY(cudaMemcpyAsync(&nruns, d_nruns, 4, cudaMemcpyDeviceToHost, es.st));
    CUDA_TRY(cudaStreamSynchronize(es.st));
    std::vector<uint32_t> uniq(nruns), counts(nruns);
    CUDA_TRY(cudaMemcpyAsync(uniq.data(), d_unique, (size_t)nruns * 4, cudaMemcpyDeviceToHost, es.st));
    CUDA_TRY(cudaMemcpyAsync(counts.data(), d_counts, (size_t)nruns * 4, cudaMemcpyDeviceToHost, es.st));
    CUDA_TRY(cudaStreamSynchronize(es.st));
    // gather list: positions (1-based, into the grouped array) of the scheduled ranks of every run
    std::vector<uint64_t> pos, rk;
    std::vector<size_t> run_begin(nruns + 1, 0);
    uint64_t off = 0;
    for (uint32_t r = 0; r < nruns; r++) {
        rank_schedule(counts[r], rk, 256, 0.0025);
        run_begin[r] = pos.size();
        for (uint64_t x : rk) pos.push_back(off + x);
        off += counts[r];
    }
    run_begin[nruns] = pos.size();
    std::vector<uint64_t> vals(pos.size());
    if (!pos.empty()) {
        uint64_t *d_pos = nullptr, *d_out = nullptr;
        CUDA_TRY(cudaMallocAsync((void**)&d_pos, pos.size() * 8, es.st)); es.temps.push_back(d_pos);
        CUDA_TRY(cudaMallocAsync((void**)&d_out, pos.size() * 8, es.st)); es.temps.push_back(d_out);
        CUDA_TRY(cudaMemcpyAsync(d_pos, pos.data(), pos.size() * 8, cudaMemcpyHostToDevice, es.st));
        k_gather_ranks<<<(unsigned)std::min<uint64_t>((pos.size() + 255) / 256, 4096), 256, 0, es.st>>>(es.pct_codes[k], d_pos, pos.size(), d_out);
        es.ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(vals.data(), d_out, pos.size() * 8, cudaMemcpyDeviceToHost, es.st));
        CUDA_TRY(cudaStreamSynchronize(es.st));
    }
    off = 0;
    for (uint32_t r = 0; r < nruns; r++) {
        PctSummary sum;
        sum.n_total = counts[r];
        for (size_t i = run_begin[r]; i < run_begin[r + 1]; i++) {
            sum.ranks.push_back(pos[i] - off);
            sum.value_bits.push_back(code_to_bits_h(TAGG_F64, vals[i]));
        }
        off += counts[r];
        // raw bucket index of the enclosing scope -> compact bucket index of the result (compact.cu rank array)
        if (uniq[r] >= rank_of_raw.size()) return tagg_fail(TAGG_ERR_CUDA, "percentile values in a bucket that does not exist (internal error)");
        res->pcts[k][(uint64_t)rank_of_raw[uniq[r]]] = std::move(sum);
    }
    return 0;
}

int read_percentiles(ExecState& es, tagg_result* res) {
    const PlanMeta& m = *es.meta;
    res->pcts.clear();
    res->pcts.resize(m.pct_node.size());
    for (size_t k = 0; k < m.pct_node.size(); k++) {
        if (es.rank[k].active) {  // rank-bin mode: collected (and checked) in exec_run
            res->pcts[k][0] = std::move(es.rank[k].summary);
            continue;
        }
        if (!es.pct_count[k]) {  // never materialised (no candidate reached the leaf's launch)
            res->pcts[k][0] = PctSummary();
            continue;
        }
        PctSummary sum;
        unsigned long long n = 0;
        CUDA_TRY(cudaMemcpyAsync(&n, es.pct_count[k], 8, cudaMemcpyDeviceToHost, es.st));
        CUDA_TRY(cudaStreamSynchronize(es.st));
        if (n > es.pct_cap[k]) n = es.pct_cap[k];
        if (m.scope_of[m.pct_node[k]] != 0) {  // under a bucket aggregation
            if (n) {
                const int sc = m.scope_of[m.pct_node[k]];
                if ((size_t)sc >= es.compact.d_rank.size() || !es.compact.d_rank[sc]) return tagg_fail(TAGG_ERR_CUDA, "no rank array for scope %d (internal error)", sc);
                std::vector<uint32_t> rank_of_raw(es.compact.cap_scope[sc]);
                CUDA_TRY(cudaMemcpyAsync(rank_of_raw.data(), es.compact.d_rank[sc], rank_of_raw.size() * 4, cudaMemcpyDeviceToHost, es.st));
                CUDA_TRY(cudaStreamSynchronize(es.st));
                int rc = read_nested_percentiles(es, res, k, n, rank_of_raw);
                if (rc) return rc;
            }
            continue;
        }
        sum.n_total = n;
        if (n > 0x7fffffffull) return tagg_fail(TAGG_ERR_UNSUPPORTED, "more than 2^31-1 percentile values in one call (%llu): split the call by segment and merge", (unsigned long long)n);
        if (n) {
            uint64_t* d_alt = nullptr;
            void* d_tmp = nullptr;
            CUDA_TRY(cudaMallocAsync((void**)&d_alt, n * 8, es.st));
            cub::DoubleBuffer<uint64_t> keys(es.pct_codes[k], d_alt);
            size_t tmp_bytes = 0;
            cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, (int)n, 0, 64, es.st);
            CUDA_TRY(cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 16, es.st));
            CUDA_TRY(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, keys, (int)n, 0, 64, es.st));
            rank_schedule(n, sum.ranks);
            uint64_t *d_ranks = nullptr, *d_out = nullptr;
            size_t nr = sum.ranks.size();
            CUDA_TRY(cudaMallocAsync((void**)&d_ranks, nr * 8, es.st));
            CUDA_TRY(cudaMallocAsync((void**)&d_out, nr * 8, es.st));
            CUDA_TRY(cudaMemcpyAsync(d_ranks, sum.ranks.data(), nr * 8, cudaMemcpyHostToDevice, es.st));
            k_gather_ranks<<<(unsigned)std::min<uint64_t>((nr + 255) / 256, 1024), 256, 0, es.st>>>(keys.Current(), d_ranks, nr, d_out);
            es.ctx->launches++;
            CUDA_TRY(cudaGetLastError());
            sum.value_bits.resize(nr);
            CUDA_TRY(cudaMemcpyAsync(sum.value_bits.data(), d_out, nr * 8, cudaMemcpyDeviceToHost, es.st));
            CUDA_TRY(cudaStreamSynchronize(es.st));
            for (auto& v : sum.value_bits) v = code_to_bits_h(TAGG_F64, v);
            cudaFreeAsync(d_alt, es.st); cudaFreeAsync(d_tmp, es.st); cudaFreeAsync(d_ranks, es.st); cudaFreeAsync(d_out, es.st);
        }
        res->pcts[k][0] = std::move(sum);
    }
    return 0;
}

// ---- the two representations of a result (host.h) ----------------------------------------------------------------
void tagg_result::materialize() {
    if (!has_img) return;
    result_ensure_host(this);
    const size_t ns = n_scope.size(), nk = off_values.size();
    scopes.resize(ns);
    slots.resize(nk);
    for (size_t s = 0; s < ns; s++) {
        const uint64_t n = n_scope[s];
        if (s == 0) { scopes[0].keys.assign(1, 0); scopes[0].parents.assign(1, 0); continue; }
        scopes[s].keys.assign(scope_keys(s), scope_keys(s) + n);
        scopes[s].parents.assign(scope_parents(s), scope_parents(s) + n);
    }
    for (size_t k = 0; k < nk; k++) {
        const uint64_t n = slot_len(k);
        slots[k].values.assign(slot_values(k), slot_values(k) + n);
        slots[k].seen.assign(slot_seen(k), slot_seen(k) + n);
    }
    has_img = false;
    release_device();  // the device image no longer describes this result
}
void tagg_result::release_device() {
    if (!d_img) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (d_stream) cudaFreeAsync(d_img, d_stream); else cudaFree(d_img);
    d_img = nullptr;
    d_bytes = 0;
}
tagg_result::~tagg_result() {
    if (d_img) { cudaFree(d_img); d_img = nullptr; }
    if (img) { if (img_pinned) cudaFreeHost(img); else free(img); img = nullptr; }
}

// ---- PreparedAgg::merge on compact results --------------------------------------------------------
static inline double bits_f64(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static inline uint64_t f64_bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }
static inline bool lt_bits(int kind, uint64_t a, uint64_t b) {
    if (kind == TAGG_U64) return a < b;
    if (kind == TAGG_F64) return bits_f64(a) < bits_f64(b);
    return (int64_t)a < (int64_t)b;
}

struct PairHash {
    size_t operator()(const std::pair<uint32_t, uint64_t>& p) const {
        uint64_t z = p.second ^ ((uint64_t)p.first * 0x9E3779B97F4A7C15ull);
        z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27;
        return (size_t)z;
    }
};

static void merge_pct(PctSummary& a, const PctSummary& b) {
    if (b.n_total == 0) return;
    if (a.n_total == 0) { a = b; return; }
    // value order for f64 bits: compare as doubles (no NaN expected in percentile inputs)
    auto less = [](uint64_t x, uint64_t y) { return bits_f64(x) < bits_f64(y); };
    bool exact = a.ranks.size() == a.n_total && b.ranks.size() == b.n_total;
    PctSummary o;
    o.n_total = a.n_total + b.n_total;
    if (exact) {  // both hold every element: an exact merge of two sorted lists
        o.value_bits.resize(o.n_total);
        std::merge(a.value_bits.begin(), a.value_bits.end(), b.value_bits.begin(), b.value_bits.end(), o.value_bits.begin(), less);
        o.ranks.resize(o.n_total);
        for (uint64_t i = 0; i < o.n_total; i++) o.ranks[i] = i + 1;
    } else {
        // rank of x in the union ~= rank_a(x) + (rank in b of the largest stored b-value <= x); the
        // error is bounded by the gap between stored ranks of the other side (<= eps/4 relative)
        struct E { uint64_t v, r; };
        std::vector<E> out;
        size_t j = 0;
        uint64_t rb = 0;
        for (size_t i = 0; i < a.ranks.size(); i++) {
            while (j < b.ranks.size() && !less(a.value_bits[i], b.value_bits[j])) rb = b.ranks[j++];
            out.push_back({a.value_bits[i], a.ranks[i] + rb});
        }
        j = 0;
        uint64_t ra = 0;
        for (size_t i = 0; i < b.ranks.size(); i++) {
            while (j < a.ranks.size() && less(a.value_bits[j], b.value_bits[i])) ra = a.ranks[j++];
            out.push_back({b.value_bits[i], b.ranks[i] + ra});
        }
        std::sort(out.begin(), out.end(), [&](const E& x, const E& y) { return x.r < y.r || (x.r == y.r && less(x.v, y.v)); });
        for (auto& e : out) {
            if (!o.ranks.empty() && o.ranks.back() == e.r) continue;
            o.ranks.push_back(e.r);
            o.value_bits.push_back(e.v);
        }
    }
    a = std::move(o);
}

int result_merge(tagg_result* dst, const tagg_result* src) {
    if (!dst->meta || !src->meta) return tagg_fail(TAGG_ERR_BAD_ARG, "result has no plan (already freed?)");
    if (dst->meta.get() != src->meta.get()) {  // different plan objects: the trees must be structurally identical
        const PlanMeta &a = *dst->meta, &b = *src->meta;
        bool same = a.nodes.size() == b.nodes.size();
        for (size_t i = 0; same && i < a.nodes.size(); i++) {
            const tagg_node &x = a.nodes[i], &y = b.nodes[i];
            same = x.op == y.op && x.kind == y.kind && x.multi == y.multi && x.field_id == y.field_id && x.n_children == y.n_children &&
                   memcmp(&x.f0, &y.f0, 8) == 0 && memcmp(&x.f1, &y.f1, 8) == 0 &&
                   (x.op != TAGG_OP_POST_FILTER || (x.pred == y.pred && x.u0 == y.u0 && x.u1 == y.u1));
        }
        if (!same) return tagg_fail(TAGG_ERR_BAD_ARG, "results come from different plans");
    }
    const PlanMeta& m = *dst->meta;
    size_t ns = m.scope_node.size();
    if (dst->merged_elsewhere || src->merged_elsewhere) return tagg_fail(TAGG_ERR_BAD_ARG, "the fruit of a tagg_execute_reduce call lives on the root rank only");
    dst->materialize();
    {
        int rce = result_ensure_host(const_cast<tagg_result*>(src));
        if (rce) return rce;
    }
    auto shaped = [&](const tagg_result* r) {
        if (r->pcts.size() != m.pct_node.size()) return false;
        if (r->has_img) return r->n_scope.size() == ns && r->off_values.size() == m.slot_node.size();
        return r->scopes.size() == ns && r->slots.size() == m.slot_node.size();
    };
    if (!shaped(dst) || !shaped(src)) return tagg_fail(TAGG_ERR_BAD_ARG, "result does not have the shape of its plan");
    std::vector<std::vector<uint32_t>> map(ns);  // src bucket -> dst bucket, per scope
    map[0] = {0};
    for (size_t s = 1; s < ns; s++) {
        auto& D = dst->scopes[s];
        const uint64_t sn = src->scope_len(s);
        const uint64_t* skeys = src->scope_keys(s);
        const uint32_t* sparents = src->scope_parents(s);
        int ps = m.scope_parent[s];
        std::unordered_map<std::pair<uint32_t, uint64_t>, uint32_t, PairHash> index;
        index.reserve(D.keys.size() * 2 + sn);
        for (size_t i = 0; i < D.keys.size(); i++) index[{D.parents[i], D.keys[i]}] = (uint32_t)i;
        map[s].resize(sn);
        for (size_t i = 0; i < sn; i++) {
            if (sparents[i] >= map[ps].size()) return tagg_fail(TAGG_ERR_BAD_ARG, "malformed result: parent bucket out of range");
            uint32_t dp = map[ps][sparents[i]];
            auto key = std::make_pair(dp, skeys[i]);
            auto it = index.find(key);
            if (it == index.end()) {  // or_insert_with(create_fruit)
                uint32_t at = (uint32_t)D.keys.size();
                D.keys.push_back(skeys[i]);
                D.parents.push_back(dp);
                index[key] = at;
                map[s][i] = at;
            } else {
                map[s][i] = it->second;
            }
        }
    }
    for (size_t k = 0; k < m.slot_node.size(); k++) {
        int node = m.slot_node[k];
        const tagg_node& nd = m.nodes[node];
        int s = m.scope_of[node];
        auto& D = dst->slots[k];
        const uint64_t sn = src->slot_len(k);
        const uint64_t* svalues = src->slot_values(k);
        const uint8_t* sseen = src->slot_seen(k);
        if (sn != map[s].size()) return tagg_fail(TAGG_ERR_BAD_ARG, "malformed result: metric length differs from its scope");
        size_t nb = dst->scopes[s].keys.size();
        D.values.resize(nb, 0);
        D.seen.resize(nb, nd.op == TAGG_OP_COUNT ? 1 : 0);
        for (size_t i = 0; i < sn; i++) {
            uint32_t d = map[s][i];
            if (nd.op == TAGG_OP_COUNT) { D.values[d] += svalues[i]; D.seen[d] = 1; continue; }
            if (!sseen[i]) continue;  // None => return
            if (!D.seen[d]) { D.values[d] = svalues[i]; D.seen[d] = 1; continue; }  // acc.replace(v)
            uint64_t v = svalues[i], &acc = D.values[d];
            if (nd.op == TAGG_OP_SUM) acc = nd.kind == TAGG_F64 ? f64_bits(bits_f64(acc) + bits_f64(v)) : acc + v;
            else if (nd.op == TAGG_OP_MIN) { if (lt_bits(nd.kind, v, acc)) acc = v; }
            else { if (lt_bits(nd.kind, acc, v)) acc = v; }
        }
    }
    for (size_t k = 0; k < m.pct_node.size(); k++) {
        const int sc = m.scope_of[m.pct_node[k]];
        for (auto& kv : src->pcts[k]) merge_pct(dst->pcts[k][sc == 0 ? kv.first : (uint64_t)map[sc][kv.first]], kv.second);
    }
    dst->kernel_ms += src->kernel_ms;
    dst->alg_bytes += src->alg_bytes;
    dst->n_launches += src->n_launches;
    return 0;
}

// ---- readers --------------------------------------------------------------------------------------
static int scope_index(const tagg_result* res, uint32_t scope_node) {
    if (scope_node == TAGG_ROOT_SCOPE) return 0;
    if (scope_node >= res->meta->nodes.size() || res->meta->own_scope[scope_node] < 0) return -1;
    return res->meta->own_scope[scope_node];
}

extern "C" {

int tagg_result_scope_len(const tagg_result* res, uint32_t scope_node, uint64_t* n_buckets) {
    if (!res || !n_buckets) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    int s = scope_index(res, scope_node);
    if (s < 0) return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a bucket aggregation", scope_node);
    *n_buckets = res->scope_len(s);
    return 0;
}

int tagg_result_scope_read(const tagg_result* res, uint32_t scope_node, uint64_t* keys, uint32_t* parents, uint64_t cap) {
    if (res) { int rce = result_ensure_host(const_cast<tagg_result*>(res)); if (rce) return rce; }
    if (!res) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    int s = scope_index(res, scope_node);
    if (s < 0) return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a bucket aggregation", scope_node);
    const uint64_t n = res->scope_len(s);
    if (cap < n) return tagg_fail(TAGG_ERR_BAD_ARG, "buffer too small");
    if (s == 0) { if (keys) keys[0] = 0; if (parents) parents[0] = 0; return 0; }
    if (keys && n) memcpy(keys, res->scope_keys(s), n * 8);
    if (parents && n) memcpy(parents, res->scope_parents(s), n * 4);
    return 0;
}

int tagg_result_metric_len(const tagg_result* res, uint32_t node, uint64_t* n_buckets) {
    if (!res || !n_buckets) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    if (node >= res->meta->nodes.size() || res->meta->slot_of[node] < 0)
        return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a count/sum/min/max leaf", node);
    *n_buckets = res->slot_len(res->meta->slot_of[node]);
    return 0;
}

int tagg_result_metric_read(const tagg_result* res, uint32_t node, uint64_t* values, uint8_t* seen, uint64_t cap) {
    if (res) { int rce = result_ensure_host(const_cast<tagg_result*>(res)); if (rce) return rce; }
    if (!res) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    if (node >= res->meta->nodes.size() || res->meta->slot_of[node] < 0)
        return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a count/sum/min/max leaf", node);
    const size_t k = res->meta->slot_of[node];
    const uint64_t n = res->slot_len(k);
    if (cap < n) return tagg_fail(TAGG_ERR_BAD_ARG, "buffer too small");
    if (values && n) memcpy(values, res->slot_values(k), n * 8);
    if (seen && n) memcpy(seen, res->slot_seen(k), n);
    return 0;
}

int tagg_result_scope_view(const tagg_result* res, uint32_t scope_node, const uint64_t** keys, const uint32_t** parents, uint64_t* n) {
    if (res) { int rce = result_ensure_host(const_cast<tagg_result*>(res)); if (rce) return rce; }
    if (!res || !n) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    int s = scope_index(res, scope_node);
    if (s <= 0) return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a bucket aggregation", scope_node);
    *n = res->scope_len(s);
    if (keys) *keys = *n ? res->scope_keys(s) : nullptr;
    if (parents) *parents = *n ? res->scope_parents(s) : nullptr;
    return 0;
}

int tagg_result_metric_view(const tagg_result* res, uint32_t node, const uint64_t** values, const uint8_t** seen, uint64_t* n) {
    if (res) { int rce = result_ensure_host(const_cast<tagg_result*>(res)); if (rce) return rce; }
    if (!res || !n) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    if (node >= res->meta->nodes.size() || res->meta->slot_of[node] < 0)
        return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a count/sum/min/max leaf", node);
    const size_t k = res->meta->slot_of[node];
    *n = res->slot_len(k);
    if (values) *values = *n ? res->slot_values(k) : nullptr;
    if (seen) *seen = *n ? res->slot_seen(k) : nullptr;
    return 0;
}

static const PctSummary* find_pct(const tagg_result* res, uint32_t node, uint64_t bucket) {
    if (node >= res->meta->nodes.size() || res->meta->pct_of[node] < 0) return nullptr;
    const auto& mp = res->pcts[res->meta->pct_of[node]];
    auto it = mp.find(bucket);
    static const PctSummary empty;
    return it == mp.end() ? &empty : &it->second;
}

int tagg_result_percentiles_len(const tagg_result* res, uint32_t node, uint64_t bucket, uint64_t* n_total, uint64_t* n_pairs) {
    if (!res) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    const PctSummary* p = find_pct(res, node, bucket);
    if (!p) return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a percentiles leaf", node);
    if (n_total) *n_total = p->n_total;
    if (n_pairs) *n_pairs = p->ranks.size();
    return 0;
}

int tagg_result_percentiles_read(const tagg_result* res, uint32_t node, uint64_t bucket, uint64_t* ranks, uint64_t* value_bits,
                                 uint64_t cap) {
    if (!res) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    const PctSummary* p = find_pct(res, node, bucket);
    if (!p) return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a percentiles leaf", node);
    if (cap < p->ranks.size()) return tagg_fail(TAGG_ERR_BAD_ARG, "buffer too small");
    if (ranks && !p->ranks.empty()) memcpy(ranks, p->ranks.data(), p->ranks.size() * 8);
    if (value_bits && !p->value_bits.empty()) memcpy(value_bits, p->value_bits.data(), p->value_bits.size() * 8);
    return 0;
}

int tagg_result_stats(const tagg_result* res, double* kernel_ms, uint64_t* alg_bytes, uint32_t* n_launches, uint32_t* path_used) {
    if (!res) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    if (kernel_ms) *kernel_ms = res->kernel_ms;
    if (alg_bytes) *alg_bytes = res->alg_bytes;
    if (n_launches) *n_launches = res->n_launches;
    if (path_used) *path_used = res->path_used;
    return 0;
}

}  // extern "C"

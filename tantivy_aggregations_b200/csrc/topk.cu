// topk.cu — `Terms::top_k(k, sort_by)` (terms.rs:425-457) and row reads on the DEVICE-resident fruit image.
//
// The reference returns the whole bucket map from agg_search and selects the top k on the host afterwards (a BinaryHeap
// over every bucket).  With the fruit image kept in HBM (compact.cu) a lazily read result (tagg_plan_set_readout) never
// ships the full table: top_k runs a radix select over the chosen leaf metric on the device and only the k winning rows
// cross PCIe ("for 1 M-bucket tables return only top-k by a payload field").
//
// Order (terms.rs:437-456): descending by the sort value; among equal sort values ascending by bucket key.  The sort value
// is the leaf metric as Rust sees it — u64 / i64 / date by value, Option<T> with None below every Some, f64 by the total
// order of the order-preserving codes (the reference needs an `Ord` key, so a caller maps f64 to an ordered type first).
#include <string.h>

#include <algorithm>

#include "exec.h"

#define TK_THREADS 1024
#define TK_MAX_TIES 65536

// sortable composite of bucket j: the metric's order-preserving code; the Option flag rides in a separate top "digit"
__device__ __forceinline__ uint64_t tk_code(uint32_t op, uint32_t kind, uint64_t bits) {
    if (op == TAGG_OP_COUNT || kind == TAGG_U64) return bits;
    if (kind == TAGG_F64) return (bits >> 63) == 0 ? bits ^ 0x8000000000000000ull : ~bits;
    return bits ^ 0x8000000000000000ull;
}

struct TkParams {
    const uint64_t* values;
    const uint8_t* seen;
    const uint32_t* parents;  // nullptr: every bucket qualifies
    uint32_t parent;
    uint32_t op, kind;
    uint64_t n, k;
    uint32_t* out_idx;        // [k + TK_MAX_TIES]
    uint64_t* out_hdr;        // [0] = buckets strictly above the threshold, [1] = ties written, [2] = ties in total, [3] = qualifying buckets
};

// One CTA: 9 digit passes of an MSD radix select (Option flag, then 8 bytes of the code) find the k-th largest sort
// value; a last pass writes the buckets above it and the ties at it.
__global__ void __launch_bounds__(TK_THREADS) k_topk(const __grid_constant__ TkParams p) {
    __shared__ unsigned long long hist[256];
    __shared__ unsigned long long s_prefix_code, s_k, s_total;
    __shared__ uint32_t s_prefix_seen, s_n_gt, s_n_tie;
    const uint32_t tid = threadIdx.x;
    auto qualifies = [&](uint64_t j) { return !p.parents || p.parents[j] == p.parent; };
    if (tid == 0) { s_prefix_code = 0; s_k = p.k; s_prefix_seen = 0; s_n_gt = 0; s_n_tie = 0; s_total = 0; }
    // pass -1: the Option flag (None < Some)
    for (int pass = -1; pass < 8; pass++) {
        for (uint32_t i = tid; i < 256; i += TK_THREADS) hist[i] = 0;
        __syncthreads();
        const uint64_t prefix = s_prefix_code;
        const uint32_t pseen = s_prefix_seen;
        const int shift = 56 - 8 * pass;
        for (uint64_t j = tid; j < p.n; j += TK_THREADS) {
            if (!qualifies(j)) continue;
            const uint32_t sn = p.op == TAGG_OP_COUNT ? 1u : (uint32_t)p.seen[j];
            if (pass < 0) { atomicAdd(&hist[sn], 1ull); continue; }
            if (sn != pseen) continue;
            const uint64_t c = sn ? tk_code(p.op, p.kind, p.values[j]) : 0ull;
            if (pass > 0 && (c >> (shift + 8)) != (prefix >> (shift + 8))) continue;
            atomicAdd(&hist[(c >> shift) & 255u], 1ull);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long k = s_k, acc = 0;
            int d = pass < 0 ? 1 : 255;
            if (pass < 0) s_total = hist[0] + hist[1];
            for (; d > 0; d--) {
                if (acc + hist[d] >= k) break;
                acc += hist[d];
            }
            s_k = k - acc;  // rank inside the chosen digit's group
            if (pass < 0) s_prefix_seen = (uint32_t)d;
            else s_prefix_code = prefix | ((unsigned long long)d << shift);
        }
        __syncthreads();
    }
    // collect: strictly above the threshold, then the ties
    const uint64_t T = s_prefix_code;
    const uint32_t Ts = s_prefix_seen;
    for (uint64_t j = tid; j < p.n; j += TK_THREADS) {
        if (!qualifies(j)) continue;
        const uint32_t sn = p.op == TAGG_OP_COUNT ? 1u : (uint32_t)p.seen[j];
        const uint64_t c = sn ? tk_code(p.op, p.kind, p.values[j]) : 0ull;
        const bool gt = sn > Ts || (sn == Ts && c > T);
        const bool eq = sn == Ts && c == T;
        if (gt) {
            const uint32_t at = atomicAdd(&s_n_gt, 1u);
            if (at < p.k) p.out_idx[at] = (uint32_t)j;
        } else if (eq) {
            const uint32_t at = atomicAdd(&s_n_tie, 1u);
            if (at < TK_MAX_TIES) p.out_idx[p.k + at] = (uint32_t)j;
        }
    }
    __syncthreads();
    if (tid == 0) {
        p.out_hdr[0] = s_n_gt;
        p.out_hdr[1] = min(s_n_tie, (uint32_t)TK_MAX_TIES);
        p.out_hdr[2] = s_n_tie;
        p.out_hdr[3] = s_total;
    }
}

template <typename T>
__global__ void k_rows(const T* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, T* __restrict__ dst) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}

// host-side order of the reference: sort value descending, key ascending
namespace {
struct Row { uint32_t idx; uint32_t seen; uint64_t code; uint64_t key_code; };
inline uint64_t code_h(uint32_t op, uint32_t kind, uint64_t bits) {
    if (op == TAGG_OP_COUNT || kind == TAGG_U64) return bits;
    if (kind == TAGG_F64) return (bits >> 63) == 0 ? bits ^ 0x8000000000000000ull : ~bits;
    return bits ^ 0x8000000000000000ull;
}
inline bool row_before(const Row& a, const Row& b) {
    if (a.seen != b.seen) return a.seen > b.seen;
    if (a.code != b.code) return a.code > b.code;
    return a.key_code < b.key_code;
}
}  // namespace

int result_ensure_host(tagg_result* res);

extern "C" {

int tagg_result_top_k(tagg_result* res, uint32_t scope_node, uint64_t parent_bucket, uint32_t by_node, uint64_t k,
                      uint32_t* out_buckets, uint64_t* n_out) {
    if (!res || !n_out || (k && !out_buckets)) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    const PlanMeta& m = *res->meta;
    if (scope_node >= m.nodes.size() || m.own_scope[scope_node] <= 0 || m.nodes[scope_node].op != TAGG_OP_TERMS)
        return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a terms aggregation", scope_node);
    const int s = m.own_scope[scope_node];
    if (by_node >= m.nodes.size() || m.slot_of[by_node] < 0 || m.scope_of[by_node] != s)
        return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a count/sum/min/max leaf directly under the buckets of node %u", by_node, scope_node);
    if (res->merged_elsewhere) return tagg_fail(TAGG_ERR_BAD_ARG, "the fruit of a tagg_execute_reduce call lives on the root rank only");
    const size_t slot = m.slot_of[by_node];
    const tagg_node& ln = m.nodes[by_node];
    const tagg_node& kn = m.nodes[scope_node];
    const bool flat = m.scope_parent[s] == 0;
    *n_out = 0;
    if (k == 0) return 0;
    const uint64_t n = res->scope_len(s);
    if (n == 0) return 0;
    std::vector<Row> rows;
    auto key_code = [&](uint64_t key_bits) { return kn.kind == TAGG_U64 ? key_bits : key_bits ^ 0x8000000000000000ull; };
    if (res->has_img && res->d_img && n > 4096 && k <= (1u << 20)) {
        // ---- on the device: radix select, then only the winners (and the ties at the cut) cross PCIe ----
        tagg_ctx* ctx = res->ctx;
        CUDA_TRY(cudaSetDevice(ctx->device));
        cudaStream_t st = res->d_stream;
        const size_t cap = k + TK_MAX_TIES;
        uint8_t* d_tmp = nullptr;
        CUDA_TRY(cudaMallocAsync((void**)&d_tmp, 64 + cap * (4 + 8 + 8 + 1) + 64, st));
        uint64_t* d_hdr = (uint64_t*)d_tmp;
        uint64_t* d_vals = (uint64_t*)(d_tmp + 64);
        uint64_t* d_keys = d_vals + cap;
        uint32_t* d_idx = (uint32_t*)(d_keys + cap);
        uint8_t* d_seen = (uint8_t*)(d_idx + cap);
        TkParams p;
        memset(&p, 0, sizeof(p));
        p.values = (const uint64_t*)(res->d_img + res->d_off_values[slot]);
        p.seen = res->d_img + res->d_off_seen[slot];
        p.parents = flat ? nullptr : (const uint32_t*)(res->d_img + res->d_off_parents[s]);
        p.parent = (uint32_t)parent_bucket;
        p.op = ln.op; p.kind = ln.kind; p.n = n; p.k = k;
        p.out_idx = d_idx; p.out_hdr = d_hdr;
        CUDA_TRY(cudaMemsetAsync(d_idx, 0, cap * 4, st));  // unused entries of the index list gather bucket 0
        k_topk<<<1, TK_THREADS, 0, st>>>(p);
        ctx->launches++;
        uint64_t hdr[4] = {0, 0, 0, 0};
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(hdr, d_hdr, 32, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { cudaFreeAsync(d_tmp, st); return tagg_fail(TAGG_ERR_CUDA, "top_k failed: %s", cudaGetErrorString(e)); }
        const uint64_t n_gt = std::min<uint64_t>(hdr[0], k), n_tie = hdr[1];
        if (hdr[2] <= TK_MAX_TIES) {
            // rows: [0, n_gt) above the threshold, [k, k + n_tie) the ties — gather value / flag / key of both runs
            std::vector<uint32_t> idx(n_gt + n_tie);
            std::vector<uint64_t> vals(idx.size()), keys(idx.size());
            std::vector<uint8_t> seen(idx.size());
            const unsigned blocks = (unsigned)std::min<uint64_t>((cap + 255) / 256, 1024);
            k_rows<uint64_t><<<blocks, 256, 0, st>>>(p.values, d_idx, cap, d_vals);
            k_rows<uint64_t><<<blocks, 256, 0, st>>>((const uint64_t*)(res->d_img + res->d_off_keys[s]), d_idx, cap, d_keys);
            k_rows<uint8_t><<<blocks, 256, 0, st>>>(p.seen, d_idx, cap, d_seen);
            ctx->launches += 3;
            auto pull = [&](void* dst, const void* src, size_t esz) {
                if (n_gt && e == cudaSuccess) e = cudaMemcpyAsync(dst, src, n_gt * esz, cudaMemcpyDeviceToHost, st);
                if (n_tie && e == cudaSuccess) e = cudaMemcpyAsync((uint8_t*)dst + n_gt * esz, (const uint8_t*)src + k * esz, n_tie * esz, cudaMemcpyDeviceToHost, st);
            };
            e = cudaGetLastError();
            pull(idx.data(), d_idx, 4);
            pull(vals.data(), d_vals, 8);
            pull(keys.data(), d_keys, 8);
            pull(seen.data(), d_seen, 1);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            cudaFreeAsync(d_tmp, st);
            if (e != cudaSuccess) return tagg_fail(TAGG_ERR_CUDA, "top_k gather failed: %s", cudaGetErrorString(e));
            rows.resize(idx.size());
            for (size_t i = 0; i < idx.size(); i++) {
                const uint32_t sn = ln.op == TAGG_OP_COUNT ? 1u : seen[i];
                rows[i] = {idx[i], sn, sn ? code_h(ln.op, ln.kind, vals[i]) : 0ull, key_code(keys[i])};
            }
            std::sort(rows.begin(), rows.end(), row_before);
            *n_out = std::min<uint64_t>(k, rows.size());
            for (uint64_t i = 0; i < *n_out; i++) out_buckets[i] = rows[i].idx;
            return 0;
        }
        cudaFreeAsync(d_tmp, st);  // a tie group larger than the scratch list at the cut: decide on the host
    }
    // ---- on the host (small scopes, merged results, huge tie groups) ----
    int rc = result_ensure_host(res);
    if (rc) return rc;
    const uint64_t* V = res->slot_values(slot);
    const uint8_t* Sn = res->slot_seen(slot);
    const uint64_t* K = res->scope_keys(s);
    const uint32_t* P = res->scope_parents(s);
    for (uint64_t j = 0; j < n; j++) {
        if (!flat && P[j] != (uint32_t)parent_bucket) continue;
        const uint32_t sn = ln.op == TAGG_OP_COUNT ? 1u : Sn[j];
        rows.push_back({(uint32_t)j, sn, sn ? code_h(ln.op, ln.kind, V[j]) : 0ull, key_code(K[j])});
    }
    const size_t kk = (size_t)std::min<uint64_t>(k, rows.size());
    std::partial_sort(rows.begin(), rows.begin() + kk, rows.end(), row_before);
    *n_out = kk;
    for (size_t i = 0; i < kk; i++) out_buckets[i] = rows[i].idx;
    return 0;
}

// Rows of a lazily read result: keys / parents of the given buckets of a scope, value / Option flag of the given buckets
// of a leaf metric — gathered on the device, only n rows cross PCIe.
int tagg_result_scope_rows(tagg_result* res, uint32_t scope_node, const uint32_t* buckets, uint64_t n, uint64_t* keys, uint32_t* parents) {
    if (!res || (n && !buckets)) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    const PlanMeta& m = *res->meta;
    if (scope_node >= m.nodes.size() || m.own_scope[scope_node] <= 0) return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a bucket aggregation", scope_node);
    const int s = m.own_scope[scope_node];
    const uint64_t len = res->scope_len(s);
    for (uint64_t i = 0; i < n; i++)
        if (buckets[i] >= len) return tagg_fail(TAGG_ERR_BAD_ARG, "bucket %u out of range (%llu buckets)", buckets[i], (unsigned long long)len);
    if (!n) return 0;
    if (!(res->lazy && res->d_img)) {
        int rc = result_ensure_host(res);
        if (rc) return rc;
        for (uint64_t i = 0; i < n; i++) { if (keys) keys[i] = res->scope_keys(s)[buckets[i]]; if (parents) parents[i] = res->scope_parents(s)[buckets[i]]; }
        return 0;
    }
    CUDA_TRY(cudaSetDevice(res->ctx->device));
    cudaStream_t st = res->d_stream;
    uint8_t* d = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&d, n * 16 + 64, st));
    uint64_t* d_keys = (uint64_t*)d;
    uint32_t* d_par = (uint32_t*)(d + n * 8);
    uint32_t* d_idx = d_par + n;
    const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, 1024);
    cudaError_t e = cudaMemcpyAsync(d_idx, buckets, n * 4, cudaMemcpyHostToDevice, st);
    k_rows<uint64_t><<<blocks, 256, 0, st>>>((const uint64_t*)(res->d_img + res->d_off_keys[s]), d_idx, n, d_keys);
    k_rows<uint32_t><<<blocks, 256, 0, st>>>((const uint32_t*)(res->d_img + res->d_off_parents[s]), d_idx, n, d_par);
    res->ctx->launches += 2;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess && keys) e = cudaMemcpyAsync(keys, d_keys, n * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && parents) e = cudaMemcpyAsync(parents, d_par, n * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(d, st);
    if (e != cudaSuccess) return tagg_fail(TAGG_ERR_CUDA, "row read failed: %s", cudaGetErrorString(e));
    return 0;
}

int tagg_result_metric_rows(tagg_result* res, uint32_t node, const uint32_t* buckets, uint64_t n, uint64_t* values, uint8_t* seen) {
    if (!res || (n && !buckets)) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    const PlanMeta& m = *res->meta;
    if (node >= m.nodes.size() || m.slot_of[node] < 0) return tagg_fail(TAGG_ERR_BAD_ARG, "node %u is not a count/sum/min/max leaf", node);
    const size_t k = m.slot_of[node];
    const uint64_t len = res->slot_len(k);
    for (uint64_t i = 0; i < n; i++)
        if (buckets[i] >= len) return tagg_fail(TAGG_ERR_BAD_ARG, "bucket %u out of range (%llu buckets)", buckets[i], (unsigned long long)len);
    if (!n) return 0;
    if (!(res->lazy && res->d_img)) {
        int rc = result_ensure_host(res);
        if (rc) return rc;
        for (uint64_t i = 0; i < n; i++) { if (values) values[i] = res->slot_values(k)[buckets[i]]; if (seen) seen[i] = res->slot_seen(k)[buckets[i]]; }
        return 0;
    }
    CUDA_TRY(cudaSetDevice(res->ctx->device));
    cudaStream_t st = res->d_stream;
    uint8_t* d = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&d, n * 16 + 64, st));
    uint64_t* d_vals = (uint64_t*)d;
    uint32_t* d_idx = (uint32_t*)(d + n * 8);
    uint8_t* d_seen = (uint8_t*)(d_idx + n);
    const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, 1024);
    cudaError_t e = cudaMemcpyAsync(d_idx, buckets, n * 4, cudaMemcpyHostToDevice, st);
    k_rows<uint64_t><<<blocks, 256, 0, st>>>((const uint64_t*)(res->d_img + res->d_off_values[k]), d_idx, n, d_vals);
    k_rows<uint8_t><<<blocks, 256, 0, st>>>(res->d_img + res->d_off_seen[k], d_idx, n, d_seen);
    res->ctx->launches += 2;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess && values) e = cudaMemcpyAsync(values, d_vals, n * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && seen) e = cudaMemcpyAsync(seen, d_seen, n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(d, st);
    if (e != cudaSuccess) return tagg_fail(TAGG_ERR_CUDA, "row read failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"

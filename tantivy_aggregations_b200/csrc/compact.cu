// compact.cu — result materialisation on the DEVICE: accumulator arena -> compact fruit image.
//
// The reference's fruits only hold buckets that a document reached (`entry().or_insert_with`, terms.rs:129-130,
// histogram.rs:148-149) and `Option` metrics.  The accumulator arena is a set of direct-indexed / hashed tables, so the
// fruit is its COMPACTION: per bucket scope the existing buckets in ascending raw order (keys decoded to the column's
// natural type, parent bucket as an index into the parent scope's order), per leaf metric the decoded value and the
// Option flag of every existing bucket.  Round 1 did this with host loops over arrays copied back one by one (1.5 ms for
// a 100 k-bucket result); here one or two small kernels per scope write a contiguous image that crosses PCIe in one
// copy, and the image stays in HBM for `top_k` on the device (terms.rs:425-457).
#include <string.h>

#include <algorithm>

#include "exec.h"

#define CP_THREADS 1024
#define CP_MAXSLOTS 16
#define CP_MAXBLOCKS 592  // 4 per SM; block counts are summed by one thread each (<= CP_THREADS)

enum { CPM_DENSE = 0, CPM_HASH = 1, CPM_ROOT = 2 };

struct CpSlot {
    const uint64_t* acc;
    const uint8_t* seen;
    uint64_t* out_values;
    uint8_t* out_seen;
    uint32_t op, kind;
};
struct CpScope {
    int32_t mode;
    uint32_t key_is_terms, key_kind, n_slots, write_scope, pad;
    uint64_t capacity, dom_min, dom_size, cells_per_block;
    const uint8_t* present;
    const uint32_t* state;
    const uint64_t* hkeys;
    const uint32_t* hparents;
    const uint32_t* parent_rank;  // compact index of every raw cell of the parent scope; nullptr: the parent is the root bucket
    uint32_t* rank;               // out (optional): compact index of every existing raw cell of this scope
    uint64_t* out_keys;
    uint32_t* out_parents;
    uint64_t* out_n;
    uint32_t* blk_count;
    CpSlot slots[CP_MAXSLOTS];
};

__device__ __forceinline__ bool cp_exists(const CpScope& sc, uint64_t i) {
    if (sc.mode == CPM_DENSE) return sc.present[i] != 0;
    if (sc.mode == CPM_HASH) return sc.state[i] == ST_READY;
    return true;
}

__global__ void __launch_bounds__(CP_THREADS) k_compact_count(const __grid_constant__ CpScope sc) {
    __shared__ uint32_t wsum[32];
    const uint64_t start = (uint64_t)blockIdx.x * sc.cells_per_block, end = min(sc.capacity, start + sc.cells_per_block);
    uint32_t c = 0;
    for (uint64_t i = start + threadIdx.x; i < end; i += CP_THREADS) c += cp_exists(sc, i) ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t t = __reduce_add_sync(0xffffffffu, wsum[threadIdx.x]);
        if (threadIdx.x == 0) sc.blk_count[blockIdx.x] = t;
    }
}

// value of a leaf metric as the ABI returns it (result readers of include/tagg.h)
__device__ __forceinline__ uint64_t cp_decode(uint32_t op, uint32_t kind, uint64_t v, bool seen) {
    if (op == TAGG_OP_COUNT) return v;
    if (!seen) return 0;
    if (op == TAGG_OP_SUM) return v;  // accumulated in the natural type already
    return code_to_bits(kind, op == TAGG_OP_MIN ? ~v : v);
}

__global__ void __launch_bounds__(CP_THREADS) k_compact_scatter(const __grid_constant__ CpScope sc) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t s_off;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t start = (uint64_t)blockIdx.x * sc.cells_per_block, end = min(sc.capacity, start + sc.cells_per_block);
    // buckets that precede this block's cells
    uint32_t off = 0;
    if (gridDim.x > 1) {
        uint32_t c = tid < blockIdx.x ? sc.blk_count[tid] : 0u;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) wsum[warp] = c;
        __syncthreads();
        if (tid < 32) {
            uint32_t t = __reduce_add_sync(0xffffffffu, wsum[tid]);
            if (tid == 0) s_off = t;
        }
        __syncthreads();
        off = s_off;
        __syncthreads();
    }
    uint32_t running = off;
    const bool one_parent = sc.dom_size == sc.capacity;
    for (uint64_t tile = start; tile < end; tile += CP_THREADS) {
        const uint64_t i = tile + tid;
        const bool ex = i < end && cp_exists(sc, i);
        const uint32_t bal = __ballot_sync(0xffffffffu, ex);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 32; w++) {
            const uint32_t x = wsum[w];
            before += w < (int)warp ? x : 0u;
            total += x;
        }
        if (ex) {
            const uint32_t j = running + before + __popc(bal & ((1u << lane) - 1u));
            if (sc.write_scope) {
                if (sc.rank) sc.rank[i] = j;
                if (sc.mode != CPM_ROOT) {
                    uint64_t key;
                    uint32_t praw;
                    if (sc.mode == CPM_DENSE) {
                        key = sc.dom_min + (one_parent ? i : i % sc.dom_size);
                        praw = one_parent ? 0u : (uint32_t)(i / sc.dom_size);
                    } else {
                        key = sc.hkeys[i];
                        praw = sc.hparents[i];
                    }
                    sc.out_keys[j] = sc.key_is_terms ? code_to_bits(sc.key_kind, key) : key;
                    sc.out_parents[j] = sc.parent_rank ? sc.parent_rank[praw] : 0u;
                }
            }
            for (uint32_t s = 0; s < sc.n_slots; s++) {
                const CpSlot& sl = sc.slots[s];
                const bool seen = sl.op == TAGG_OP_COUNT || sl.seen[i] != 0;
                sl.out_values[j] = cp_decode(sl.op, sl.kind, sl.acc[i], seen);
                sl.out_seen[j] = seen ? 1 : 0;
            }
        }
        running += total;
        __syncthreads();
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0 && sc.write_scope) *sc.out_n = running;
}

static inline size_t al64(size_t x) { return (x + 63) & ~(size_t)63; }

// Kernels: arena -> device image.  Call after the pass (and after the cross-GPU merge of the arena, if any).
int compact_launch(ExecState& es) {
    const PlanMeta& m = *es.meta;
    CompactState& C = es.compact;
    const size_t ns = es.scopes.size(), nk = es.slots.size();
    C = CompactState();
    C.cap_scope.resize(ns);
    C.d_off_keys.assign(ns, 0); C.d_off_parents.assign(ns, 0);
    C.d_off_values.assign(nk, 0); C.d_off_seen.assign(nk, 0);
    C.d_rank.assign(ns, nullptr);
    size_t off = al64(ns * 8);
    for (size_t s = 0; s < ns; s++) {
        C.cap_scope[s] = s == 0 ? 1 : es.scopes[s].capacity;
        if (s == 0) continue;
        C.d_off_keys[s] = off; off = al64(off + C.cap_scope[s] * 8);
        C.d_off_parents[s] = off; off = al64(off + C.cap_scope[s] * 4);
    }
    for (size_t k = 0; k < nk; k++) {
        const uint64_t cap = C.cap_scope[m.scope_of[m.slot_node[k]]];
        C.d_off_values[k] = off; off = al64(off + cap * 8);
        C.d_off_seen[k] = off; off = al64(off + cap);
    }
    C.d_bytes = off;
    // which scopes must publish their rank array: parents of other scopes, scopes that hold nested percentiles
    std::vector<uint8_t> need_rank(ns, 0);
    for (size_t s = 1; s < ns; s++)
        if (m.scope_parent[s] > 0) need_rank[m.scope_parent[s]] = 1;
    for (size_t k = 0; k < m.pct_node.size(); k++)
        if (m.scope_of[m.pct_node[k]] > 0) need_rank[m.scope_of[m.pct_node[k]]] = 1;
    size_t aux = 0;
    std::vector<size_t> off_rank(ns, 0), off_blk(ns, 0);
    for (size_t s = 1; s < ns; s++) {
        if (need_rank[s]) { off_rank[s] = aux; aux = al64(aux + C.cap_scope[s] * 4); }
        off_blk[s] = aux; aux = al64(aux + CP_MAXBLOCKS * 4);
    }
    void* p = nullptr;
    CUDA_TRY(cudaMallocAsync(&p, C.d_bytes + aux + 64, es.st));
    C.d_img = (uint8_t*)p;
    uint8_t* d_aux = C.d_img + C.d_bytes;
    CUDA_TRY(cudaMemsetAsync(C.d_img, 0, al64(ns * 8), es.st));  // header: buckets per scope

    for (size_t s = 0; s < ns; s++) {
        std::vector<size_t> slots_here;
        for (size_t k = 0; k < nk; k++)
            if ((size_t)m.scope_of[m.slot_node[k]] == s) slots_here.push_back(k);
        if (s == 0 && slots_here.empty()) continue;
        CpScope sc;
        memset(&sc, 0, sizeof(sc));
        const ScopeLayout& L = es.scopes[s];
        sc.capacity = C.cap_scope[s];
        if (s == 0) {
            sc.mode = CPM_ROOT;
            sc.dom_size = 1;
        } else {
            const tagg_node& nd = m.nodes[m.scope_node[s]];
            sc.mode = L.mode == SCOPE_DENSE ? CPM_DENSE : CPM_HASH;
            sc.key_is_terms = nd.op == TAGG_OP_TERMS;
            sc.key_kind = nd.kind;
            sc.dom_min = L.dom_min;
            sc.dom_size = L.mode == SCOPE_DENSE ? L.dom_size : 0;
            if (L.mode == SCOPE_DENSE) {
                sc.present = es.arena + L.off_present;
            } else {
                sc.state = (const uint32_t*)(es.arena + L.off_state);
                sc.hkeys = (const uint64_t*)(es.arena + L.off_keys);
                sc.hparents = (const uint32_t*)(es.arena + L.off_parents);
            }
            const int ps = m.scope_parent[s];
            sc.parent_rank = ps > 0 ? C.d_rank[ps] : nullptr;
            if (need_rank[s]) { C.d_rank[s] = (uint32_t*)(d_aux + off_rank[s]); sc.rank = C.d_rank[s]; }
            sc.out_keys = (uint64_t*)(C.d_img + C.d_off_keys[s]);
            sc.out_parents = (uint32_t*)(C.d_img + C.d_off_parents[s]);
            sc.blk_count = (uint32_t*)(d_aux + off_blk[s]);
        }
        sc.out_n = (uint64_t*)C.d_img + s;
        uint64_t blocks = std::min<uint64_t>((sc.capacity + CP_THREADS * 8 - 1) / (CP_THREADS * 8), CP_MAXBLOCKS);
        if (blocks < 1) blocks = 1;
        sc.cells_per_block = ((sc.capacity + blocks - 1) / blocks + CP_THREADS - 1) / CP_THREADS * CP_THREADS;
        if (sc.cells_per_block == 0) sc.cells_per_block = CP_THREADS;
        blocks = std::max<uint64_t>(1, (sc.capacity + sc.cells_per_block - 1) / sc.cells_per_block);
        if (blocks > 1) {
            k_compact_count<<<(unsigned)blocks, CP_THREADS, 0, es.st>>>(sc);
            CUDA_TRY(cudaGetLastError());
            es.ctx->launches++;
            es.n_launches++;
        }
        // the scope's own arrays ride with the first batch of slots
        for (size_t at = 0; at == 0 || at < slots_here.size(); at += CP_MAXSLOTS) {
            sc.write_scope = at == 0;
            sc.n_slots = (uint32_t)std::min<size_t>(CP_MAXSLOTS, slots_here.size() - at);
            for (uint32_t i = 0; i < sc.n_slots; i++) {
                const size_t k = slots_here[at + i];
                const tagg_node& ln = m.nodes[m.slot_node[k]];
                CpSlot& sl = sc.slots[i];
                sl.acc = (const uint64_t*)(es.arena + es.slots[k].off_acc);
                sl.seen = es.arena + es.slots[k].off_seen;
                sl.out_values = (uint64_t*)(C.d_img + C.d_off_values[k]);
                sl.out_seen = C.d_img + C.d_off_seen[k];
                sl.op = ln.op;
                sl.kind = ln.kind;
            }
            k_compact_scatter<<<(unsigned)blocks, CP_THREADS, 0, es.st>>>(sc);
            CUDA_TRY(cudaGetLastError());
            es.ctx->launches++;
            es.n_launches++;
        }
    }
    C.launched = true;
    return 0;
}

static int result_reserve(tagg_result* res, size_t bytes) {
    if (res->img && res->img_cap >= bytes) return 0;
    if (res->img) { if (res->img_pinned) cudaFreeHost(res->img); else free(res->img); }
    res->img = nullptr;
    size_t cap = 1 << 16;
    while (cap < bytes) cap <<= 1;
    void* p = nullptr;
    if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) == cudaSuccess) {
        res->img_pinned = true;
    } else {
        cudaGetLastError();
        p = malloc(cap);
        if (!p) return tagg_fail(TAGG_ERR_OOM, "result image allocation failed (%zu bytes)", cap);
        res->img_pinned = false;
    }
    res->img = (uint8_t*)p;
    res->img_cap = cap;
    return 0;
}

// Issues the download: small images whole (one copy, valid after the caller's sync), else only the header.
int compact_download_begin(ExecState& es, tagg_result* res) {
    CompactState& C = es.compact;
    const size_t ns = es.scopes.size();
    C.lazy = es.plan->readout == TAGG_READOUT_LAZY;
    C.one_shot = !C.lazy && C.d_bytes <= ((size_t)8 << 20);
    const size_t bytes = C.one_shot ? C.d_bytes : al64(ns * 8);
    int rc = result_reserve(res, bytes);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(res->img, C.d_img, bytes, cudaMemcpyDeviceToHost, es.st));
    return 0;
}

// After the sync: exact-size copies when the image was too large for one shot, then the result's directory.
int compact_finish(ExecState& es, tagg_result* res) {
    CompactState& C = es.compact;
    const size_t ns = es.scopes.size(), nk = es.slots.size();
    res->n_scope.assign(ns, 0);
    memcpy(res->n_scope.data(), res->img, ns * 8);
    res->n_scope[0] = 1;
    res->off_keys.assign(ns, 0); res->off_parents.assign(ns, 0);
    res->off_values.assign(nk, 0); res->off_seen.assign(nk, 0);
    for (size_t s = 1; s < ns; s++)
        if (res->n_scope[s] > C.cap_scope[s]) return tagg_fail(TAGG_ERR_CUDA, "compaction produced more buckets than cells (internal error)");
    // the device image moves into the result: top_k / row reads run on it (released with the result)
    res->release_device();
    res->d_img = C.d_img;
    res->d_bytes = C.d_bytes;
    res->d_off_keys = C.d_off_keys;
    res->d_off_parents = C.d_off_parents;
    res->d_off_values = C.d_off_values;
    res->d_off_seen = C.d_off_seen;
    C.d_img = nullptr;
    res->has_img = true;
    res->lazy = false;
    for (auto& sc : res->scopes) { sc.keys.clear(); sc.parents.clear(); }
    for (auto& sl : res->slots) { sl.values.clear(); sl.seen.clear(); }
    if (C.lazy) {  // only the bucket counts came down; the arrays follow on demand (result_ensure_host / row reads)
        res->lazy = true;
        return 0;
    }
    if (C.one_shot) {
        res->off_keys = C.d_off_keys; res->off_parents = C.d_off_parents;
        res->off_values = C.d_off_values; res->off_seen = C.d_off_seen;
        return 0;
    }
    res->lazy = true;
    return result_ensure_host(res);
}

// exact-size copies of every array of the device image into the result's host image
int result_ensure_host(tagg_result* res) {
    if (!res->has_img || !res->lazy) return 0;
    if (!res->d_img) return tagg_fail(TAGG_ERR_BAD_ARG, "the device image of this result is gone");
    const PlanMeta& m = *res->meta;
    const size_t ns = res->n_scope.size(), nk = res->d_off_values.size();
    CUDA_TRY(cudaSetDevice(res->ctx->device));
    cudaStream_t st = res->d_stream;
    {
        const uint8_t* d_img = res->d_img;
        const struct { const std::vector<size_t>&k, &p, &v, &s; } D{res->d_off_keys, res->d_off_parents, res->d_off_values, res->d_off_seen};
        res->off_keys.assign(ns, 0); res->off_parents.assign(ns, 0);
        res->off_values.assign(nk, 0); res->off_seen.assign(nk, 0);
        size_t off = al64(ns * 8);
        for (size_t s = 1; s < ns; s++) {
            res->off_keys[s] = off; off = al64(off + res->n_scope[s] * 8);
            res->off_parents[s] = off; off = al64(off + res->n_scope[s] * 4);
        }
        for (size_t k = 0; k < nk; k++) {
            const uint64_t n = res->n_scope[m.scope_of[m.slot_node[k]]];
            res->off_values[k] = off; off = al64(off + n * 8);
            res->off_seen[k] = off; off = al64(off + n);
        }
        int rc = result_reserve(res, off);
        if (rc) return rc;
        for (size_t s = 1; s < ns; s++) {
            const uint64_t n = res->n_scope[s];
            if (!n) continue;
            CUDA_TRY(cudaMemcpyAsync(res->img + res->off_keys[s], d_img + D.k[s], n * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(res->img + res->off_parents[s], d_img + D.p[s], n * 4, cudaMemcpyDeviceToHost, st));
        }
        for (size_t k = 0; k < nk; k++) {
            const uint64_t n = res->n_scope[m.scope_of[m.slot_node[k]]];
            if (!n) continue;
            CUDA_TRY(cudaMemcpyAsync(res->img + res->off_values[k], d_img + D.v[k], n * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(res->img + res->off_seen[k], d_img + D.s[k], n, cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    res->lazy = false;
    return 0;
}

void compact_release(ExecState& es) {
    if (es.compact.d_img) cudaFreeAsync(es.compact.d_img, es.st);
    es.compact = CompactState();
}

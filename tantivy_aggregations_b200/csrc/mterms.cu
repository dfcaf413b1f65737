// mterms.cu — K5: terms buckets keyed by a MULTI-VALUED field (terms.rs:172-179) or by a key domain too
// wide for a dense table (global open-addressing spill table), with count / sum / min / max leaves on
// single- or multi-valued columns (sum.rs:131-140, minmax.rs:135-145) — BASELINE config C4.
//
// The reference touches a bucket once per VALUE OCCURRENCE of the key field and lets the nested leaves
// collect the DOCUMENT each time, i.e. sum_agg_f64s under terms_agg_u64s adds all of the document's
// values once per key occurrence.  Here a sub-block of 256 threads owns a tile of 1024 documents:
//   doc phase    one thread per document: docset / deletes / predicates, the document's key range
//                [idx[d], idx[d+1]) and its leaf contribution folded ONCE (sum / min / max over the
//                document's values) into shared memory;
//   expand       every document writes its tile-local index over its key range (u16 per key occurrence);
//   value phase  one thread per KEY OCCURRENCE: consecutive lanes unpack consecutive packed keys (fully
//                coalesced, each 32-byte sector of the key column is read once), find the bucket (dense:
//                direct index, hashed: CAS-claimed open addressing) and apply the document's folded
//                contribution with one RED per leaf.
// The bucket tables live in global memory (L2): with 10^6 uniformly hit keys nothing smaller than the table
// itself captures any reuse, so the floor of this kernel is the L2 atomic unit, not HBM
// (profiles/r1_atom_bench_b200.txt: 190 G RED.F64/s on 10^6 addresses).  Bucket-existence / Option flags
// are filtered through a CTA-private bitmap in shared memory so an occurrence costs no global load.
#include <string.h>

#include <algorithm>

#include "exec.h"

#define MT_SUB_THREADS 256
#define MT_MAXSUB 3
#define MT_TILE 1024     // documents per sub-block tile
#define MT_CHUNK 8192    // key occurrences expanded at a time
#define MT_MAXGROUPS 2
#define MT_MAXPRED 4
#define MT_MAXCOUNTS 2
#define MT_U 4           // key occurrences in flight per thread

enum { MO_SUM = 1, MO_MIN = 2, MO_MAX = 4 };
enum { MP_FILTER = 0, MP_RANGE = 1, MP_LUT = 2, MP_RANGE_ANY = 3, MP_LUT_ANY = 4 };

struct MGroup {
    int32_t col;      // device column slot (multi: idx column, values at col + 1)
    uint32_t kind, multi, ops;
    uint64_t *acc_sum, *acc_min, *acc_max;
    uint8_t* seen;    // Option flags of the group's slots (one array, aliased by all of them)
    uint32_t soff_sum, soff_min, soff_max;  // per-document folded contribution, offsets inside a sub-block's shared block
};
struct MPred {
    int32_t type, col, filter, pad;
    uint64_t lo, hi;
    const uint8_t* lut;
};
struct MParams {
    const DevSegment* segs;       // all segments of the call
    const uint32_t* tile_begin;   // n_segs + 1 tile offsets
    uint32_t n_segs, n_tiles;
    int32_t key_col, key_multi;
    DevScope scope;
    uint32_t* overflow;
    int32_t n_preds;
    MPred preds[MT_MAXPRED];
    int32_t n_counts;
    uint64_t* count_acc[MT_MAXCOUNTS];
    int32_t n_groups;
    MGroup groups[MT_MAXGROUPS];
    uint32_t full_mask;      // flag bits of a document that contributes to every group
    uint32_t bitmap_bytes;   // CTA-private "bucket fully flagged" bitmap (dense scopes), 0 = none
    uint32_t sub_bytes;      // shared bytes per sub-block
    uint32_t soff_koff, soff_docof, soff_flags;
};

__device__ __forceinline__ bool mpred_value(const MPred& pr, uint64_t code) {
    if (pr.type == MP_RANGE || pr.type == MP_RANGE_ANY) return code >= pr.lo && code <= pr.hi;
    if (code < pr.lo) return false;
    uint64_t r = code - pr.lo;
    return r < pr.hi && ((pr.lut[r >> 3] >> (r & 7)) & 1);
}

__device__ __forceinline__ void named_bar(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// the key column of the current segment, held in registers; both words of a value are always fetched (the
// allocation is padded, dev.cuh), so the two loads are independent and there is no branch on the straddle
struct KeyCol {
    const uint64_t* words;
    uint64_t minv, mask;
    uint32_t nb;
};
__device__ __forceinline__ uint64_t key_get(const KeyCol& c, uint64_t i) {
    const uint64_t bit = i * c.nb;
    const uint64_t w = bit >> 6;
    const uint32_t sh = (uint32_t)bit & 63u;
    const uint64_t lo = __ldg(c.words + w), hi = __ldg(c.words + w + 1);
    const uint64_t v = (lo >> sh) | ((hi << 1) << (63u - sh));
    return (v & c.mask) + c.minv;
}

__global__ void __launch_bounds__(MT_SUB_THREADS * MT_MAXSUB, 1) k_mterms(const __grid_constant__ MParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, sub = tid / MT_SUB_THREADS, st = tid % MT_SUB_THREADS;
    const uint32_t n_sub = blockDim.x / MT_SUB_THREADS;
    uint32_t* bitmap = (uint32_t*)smem;
    uint8_t* base = smem + p.bitmap_bytes + (size_t)sub * p.sub_bytes;
    uint32_t* koff = (uint32_t*)(base + p.soff_koff);    // MT_TILE + 1 key offsets relative to the tile's first key
    uint16_t* docof = (uint16_t*)(base + p.soff_docof);  // MT_CHUNK tile-local document indices
    uint8_t* flags = base + p.soff_flags;                // per document: bit 0 matched, bit 1 + g contributes to group g
    if (p.bitmap_bytes) {
        for (uint32_t i = tid; i < p.bitmap_bytes / 4; i += blockDim.x) bitmap[i] = 0;
        __syncthreads();
    }
    const bool dense = p.scope.mode == SCOPE_DENSE;
    uint32_t cur_seg = 0;

    for (uint64_t tile = (uint64_t)blockIdx.x * n_sub + sub; tile < p.n_tiles; tile += (uint64_t)gridDim.x * n_sub) {
        while (cur_seg + 1 < p.n_segs && __ldg(p.tile_begin + cur_seg + 1) <= (uint32_t)tile) cur_seg++;
        const DevSegment& S = p.segs[cur_seg];
        const uint32_t d0 = ((uint32_t)tile - __ldg(p.tile_begin + cur_seg)) * MT_TILE;
        const uint32_t max_doc = S.max_doc;
        const uint32_t nd = min((uint32_t)MT_TILE, max_doc - d0);
        const DevColumn& kidx = S.cols[p.key_col];
        KeyCol kc;
        {
            const DevColumn& kv = S.cols[p.key_multi ? p.key_col + 1 : p.key_col];
            kc.words = kv.words; kc.minv = kv.min_value; kc.mask = kv.mask; kc.nb = kv.num_bits;
        }
        const uint64_t kbase = p.key_multi ? col_get(kidx, d0) : (uint64_t)d0;
        // ---- doc phase ---------------------------------------------------------------------------------
        for (uint32_t i = st; i <= nd; i += MT_SUB_THREADS) {
            const uint32_t doc = d0 + i;
            koff[i] = p.key_multi ? (uint32_t)(col_get(kidx, doc) - kbase) : i;
            if (i == nd) break;
            bool ok = docset_test(S, S.main, doc);
            if (ok && S.has_deletes) ok = !((S.deleted[doc >> 5] >> (doc & 31)) & 1u);  // searcher.rs:41-46
            for (int k = 0; ok && k < p.n_preds; k++) {
                const MPred& pr = p.preds[k];
                if (pr.type == MP_FILTER) {  // filter.rs:100-122
                    ok = docset_test(S, S.filters[pr.filter], doc);
                } else if (pr.type == MP_RANGE || pr.type == MP_LUT) {  // post_filter.rs:245-249
                    ok = mpred_value(pr, col_get(S.cols[pr.col], doc));
                } else {  // post_filter.rs:289-297: any value passes
                    uint64_t a = col_get(S.cols[pr.col], doc), e = col_get(S.cols[pr.col], (uint64_t)doc + 1);
                    bool any = false;
                    for (uint64_t j = a; j < e && !any; j++) any = mpred_value(pr, col_get(S.cols[pr.col + 1], j));
                    ok = any;
                }
            }
            uint32_t f = ok ? 1u : 0u;
            if (ok) {
#pragma unroll
                for (int g = 0; g < MT_MAXGROUPS; g++) {
                    if (g >= p.n_groups) break;
                    const MGroup& G = p.groups[g];
                    uint64_t a = doc, e = (uint64_t)doc + 1;
                    if (G.multi) { a = col_get(S.cols[G.col], doc); e = col_get(S.cols[G.col], (uint64_t)doc + 1); }
                    const DevColumn& vc = S.cols[G.multi ? G.col + 1 : G.col];
                    uint64_t sum = 0, mn = 0, mx = 0;  // min in max-form (~code), like the arena
                    for (uint64_t j = a; j < e; j++) {
                        uint64_t code = col_get(vc, j);
                        if (G.ops & MO_SUM) {
                            if (G.kind == TAGG_F64) sum = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)sum), code_to_f64(code)));
                            else sum += code_to_bits(G.kind, code);
                        }
                        mn = max(mn, ~code);
                        mx = max(mx, code);
                    }
                    if (e > a) {
                        f |= 2u << g;
                        if (G.ops & MO_SUM) ((uint64_t*)(base + G.soff_sum))[i] = sum;
                        if (G.ops & MO_MIN) ((uint64_t*)(base + G.soff_min))[i] = mn;
                        if (G.ops & MO_MAX) ((uint64_t*)(base + G.soff_max))[i] = mx;
                    }
                }
            }
            flags[i] = (uint8_t)f;
        }
        named_bar(1 + sub, MT_SUB_THREADS);
        const uint32_t nk = koff[nd];
        for (uint32_t cbase = 0; cbase < nk; cbase += MT_CHUNK) {
            const uint32_t cn = min((uint32_t)MT_CHUNK, nk - cbase);
            // ---- expand: tile-local document index of every key occurrence of the chunk ------------------
            for (uint32_t i = st; i < nd; i += MT_SUB_THREADS) {
                uint32_t lo = max(koff[i], cbase), hi = min(koff[i + 1], cbase + cn);
                for (uint32_t v = lo; v < hi; v++) docof[v - cbase] = (uint16_t)i;
            }
            named_bar(1 + sub, MT_SUB_THREADS);
            // ---- value phase: one thread per key occurrence (terms.rs:172-179), MT_U occurrences in flight ---
            for (uint32_t v0 = st; v0 < cn; v0 += MT_SUB_THREADS * MT_U) {
                uint32_t di[MT_U], f[MT_U], b[MT_U];
                uint64_t key[MT_U];
#pragma unroll
                for (int u = 0; u < MT_U; u++) {
                    const uint32_t v = v0 + u * MT_SUB_THREADS;
                    di[u] = 0; f[u] = 0;
                    if (v < cn) { di[u] = docof[v]; f[u] = flags[di[u]]; }
                }
#pragma unroll
                for (int u = 0; u < MT_U; u++) {
                    key[u] = 0;
                    if (f[u]) key[u] = key_get(kc, kbase + cbase + v0 + u * MT_SUB_THREADS);
                }
#pragma unroll
                for (int u = 0; u < MT_U; u++) {
                    b[u] = INVALID_BUCKET;
                    if (!f[u]) continue;
                    if (dense) {
                        const uint64_t rel = key[u] - p.scope.dom_min;
                        if (key[u] < p.scope.dom_min || rel >= p.scope.dom_size) { f[u] = 0; continue; }
                        b[u] = (uint32_t)rel;
                        // existence / Option flags: the CTA's bitmap remembers buckets whose flags are all set
                        bool known = false;
                        if (p.bitmap_bytes) known = (bitmap[b[u] >> 5] >> (b[u] & 31)) & 1u;
                        if (!known) {
                            if (!p.scope.present[b[u]]) p.scope.present[b[u]] = 1;
                            for (int g = 0; g < p.n_groups; g++)
                                if (((f[u] >> (1 + g)) & 1u) && p.groups[g].multi && !p.groups[g].seen[b[u]]) p.groups[g].seen[b[u]] = 1;
                            if (p.bitmap_bytes && f[u] == p.full_mask) atomicOr(bitmap + (b[u] >> 5), 1u << (b[u] & 31));
                        }
                    } else {
                        b[u] = scope_lookup(p.overflow, p.scope, 0, key[u]);
                        if (b[u] == INVALID_BUCKET) { f[u] = 0; continue; }
                        for (int g = 0; g < p.n_groups; g++)
                            if (((f[u] >> (1 + g)) & 1u) && !p.groups[g].seen[b[u]]) p.groups[g].seen[b[u]] = 1;
                    }
                }
#pragma unroll
                for (int u = 0; u < MT_U; u++) {
                    if (!f[u]) continue;
                    if (p.n_counts > 0) atomicAdd((unsigned long long*)(p.count_acc[0] + b[u]), 1ull);
                    if (p.n_counts > 1) atomicAdd((unsigned long long*)(p.count_acc[1] + b[u]), 1ull);
#pragma unroll
                    for (int g = 0; g < MT_MAXGROUPS; g++) {
                        if (g >= p.n_groups || !((f[u] >> (1 + g)) & 1u)) continue;
                        const MGroup& G = p.groups[g];
                        if (G.ops & MO_SUM) {
                            const uint64_t s = ((const uint64_t*)(base + G.soff_sum))[di[u]];
                            if (G.kind == TAGG_F64) atomicAdd((double*)(G.acc_sum + b[u]), __longlong_as_double((long long)s));
                            else atomicAdd((unsigned long long*)(G.acc_sum + b[u]), (unsigned long long)s);
                        }
                        if (G.ops & MO_MIN) {
                            const uint64_t m = ((const uint64_t*)(base + G.soff_min))[di[u]];
                            if (G.acc_min[b[u]] < m && __ldcg(G.acc_min + b[u]) < m) atomicMax((unsigned long long*)(G.acc_min + b[u]), (unsigned long long)m);
                        }
                        if (G.ops & MO_MAX) {
                            const uint64_t m = ((const uint64_t*)(base + G.soff_max))[di[u]];
                            if (G.acc_max[b[u]] < m && __ldcg(G.acc_max + b[u]) < m) atomicMax((unsigned long long*)(G.acc_max + b[u]), (unsigned long long)m);
                        }
                    }
                }
            }
            named_bar(1 + sub, MT_SUB_THREADS);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host: [filter_agg | post_filter_agg_*]* -> (..., TERMS member, ...) where the TERMS node was not taken by a
// streaming launch (multi-valued key, hashed scope, multi-valued leaves) and its sub-tree is count / sum / min /
// max leaves.  One launch per segment and TERMS member; the members covered are flagged in es.skip.
// Returns the number of members handled, < 0 on error.
// ------------------------------------------------------------------------------------------------------
int mterms_try(ExecState& es) {
    const PlanMeta& m = *es.meta;
    if (es.segs.empty()) return 0;
    const uint32_t n_nodes = (uint32_t)m.nodes.size();
    for (auto& hs : es.hsegs)
        if (hs.main.kind == DS_IDS) return 0;  // sparse id lists: the gather (generic) kernel
    if (es.skip.size() != n_nodes) es.skip.assign(n_nodes, 0);

    MParams base;
    memset(&base, 0, sizeof(base));
    uint32_t node = 0;
    while (node < n_nodes && (m.nodes[node].op == TAGG_OP_FILTER || m.nodes[node].op == TAGG_OP_POST_FILTER)) {
        const tagg_node& nd = m.nodes[node];
        if (base.n_preds >= MT_MAXPRED) return 0;
        MPred& pr = base.preds[base.n_preds++];
        if (nd.op == TAGG_OP_FILTER) {
            pr.type = MP_FILTER;
            pr.filter = (int32_t)nd.aux;
        } else {
            const bool lut = nd.pred == TAGG_PRED_LUT;
            pr.type = nd.multi ? (lut ? MP_LUT_ANY : MP_RANGE_ANY) : (lut ? MP_LUT : MP_RANGE);
            pr.col = m.col_slot[node];
            pr.lo = nd.u0;
            pr.hi = nd.u1;
            pr.lut = lut ? es.plan->d_blobs[nd.aux] : nullptr;
        }
        node++;
    }
    if (node >= n_nodes) return 0;
    std::vector<uint32_t> members;
    if (m.nodes[node].op == TAGG_OP_TUPLE) {
        for (uint32_t c = node + 1; c < m.end[node]; c = m.end[c]) members.push_back(c);
    } else {
        members.push_back(node);
    }

    int handled = 0;
    for (uint32_t mem : members) {
        if (es.skip[mem]) continue;
        const tagg_node& nd = m.nodes[mem];
        if (nd.op != TAGG_OP_TERMS) continue;
        const int sc = m.own_scope[mem];
        if (m.scope_parent[sc] != 0) continue;
        const ScopeLayout& L = es.scopes[sc];
        if (L.capacity > 0xfffffff0ull) continue;
        // leaves
        std::vector<uint32_t> leaves;
        const uint32_t sub = mem + 1;
        if (m.nodes[sub].op == TAGG_OP_TUPLE) {
            for (uint32_t c = sub + 1; c < m.end[sub]; c = m.end[c]) leaves.push_back(c);
        } else {
            leaves.push_back(sub);
        }
        MParams p = base;
        bool ok = true;
        std::vector<std::pair<int, int>> alias;  // (slot, group): Option flags of the slot live in the group's array
        for (uint32_t lf : leaves) {
            const tagg_node& ln = m.nodes[lf];
            const SlotLayout& SL = es.slots[m.slot_of[lf] < 0 ? 0 : m.slot_of[lf]];
            if (ln.op == TAGG_OP_COUNT) {
                if (p.n_counts >= MT_MAXCOUNTS) { ok = false; break; }
                p.count_acc[p.n_counts++] = (uint64_t*)(es.arena + SL.off_acc);
            } else if (ln.op == TAGG_OP_SUM || ln.op == TAGG_OP_MIN || ln.op == TAGG_OP_MAX) {
                const uint32_t bit = ln.op == TAGG_OP_SUM ? MO_SUM : ln.op == TAGG_OP_MIN ? MO_MIN : MO_MAX;
                MGroup* G = nullptr;
                int gi = -1;
                for (int g = 0; g < p.n_groups; g++)
                    if (p.groups[g].col == m.col_slot[lf] && p.groups[g].kind == ln.kind && p.groups[g].multi == (ln.multi ? 1u : 0u)) { G = &p.groups[g]; gi = g; }
                if (!G) {
                    if (p.n_groups >= MT_MAXGROUPS) { ok = false; break; }
                    gi = p.n_groups++;
                    G = &p.groups[gi];
                    memset(G, 0, sizeof(*G));
                    G->col = m.col_slot[lf];
                    G->kind = ln.kind;
                    G->multi = ln.multi ? 1 : 0;
                }
                if (G->ops & bit) { ok = false; break; }
                G->ops |= bit;
                uint64_t* acc = (uint64_t*)(es.arena + SL.off_acc);
                if (bit == MO_SUM) G->acc_sum = acc;
                else if (bit == MO_MIN) G->acc_min = acc;
                else G->acc_max = acc;
                alias.push_back({m.slot_of[lf], gi});
            } else {
                ok = false;
                break;
            }
        }
        if (!ok) continue;

        // Option flags: single-valued leaves under a dense scope are Some exactly where the bucket exists;
        // otherwise one flag array per column group (the first slot's), shared by the group's slots
        const bool dense = L.mode == SCOPE_DENSE;
        std::vector<size_t> group_seen(p.n_groups, (size_t)-1);
        for (auto& a : alias) {
            SlotLayout& SL = es.slots[a.first];
            if (dense && !p.groups[a.second].multi) SL.off_seen = L.off_present;
            else if (group_seen[a.second] == (size_t)-1) group_seen[a.second] = SL.off_seen;
            else SL.off_seen = group_seen[a.second];
        }
        for (int g = 0; g < p.n_groups; g++)
            p.groups[g].seen = es.arena + (group_seen[g] == (size_t)-1 ? L.off_present : group_seen[g]);

        p.key_col = m.col_slot[mem];
        p.key_multi = nd.multi ? 1 : 0;
        p.scope.mode = L.mode;
        p.scope.parent = 0;
        p.scope.capacity = L.capacity;
        p.scope.dom_min = L.dom_min;
        p.scope.dom_size = L.dom_size;
        if (dense) {
            p.scope.present = es.arena + L.off_present;
        } else {
            p.scope.keys = (uint64_t*)(es.arena + L.off_keys);
            p.scope.parents = (uint32_t*)(es.arena + L.off_parents);
            p.scope.state = (uint32_t*)(es.arena + L.off_state);
            p.scope.used = (unsigned long long*)(es.arena + L.off_used);
        }
        p.overflow = (uint32_t*)(es.arena + es.off_overflow);
        p.full_mask = 1u;
        for (int g = 0; g < p.n_groups; g++) p.full_mask |= 2u << g;

        // shared-memory plan
        uint32_t off = 0;
        p.soff_koff = off; off += (MT_TILE + 1) * 4; off = (off + 15) & ~15u;
        p.soff_docof = off; off += MT_CHUNK * 2;
        p.soff_flags = off; off += MT_TILE; off = (off + 15) & ~15u;
        for (int g = 0; g < p.n_groups; g++) {
            MGroup& G = p.groups[g];
            if (G.ops & MO_SUM) { G.soff_sum = off; off += MT_TILE * 8; }
            if (G.ops & MO_MIN) { G.soff_min = off; off += MT_TILE * 8; }
            if (G.ops & MO_MAX) { G.soff_max = off; off += MT_TILE * 8; }
        }
        p.sub_bytes = off;
        const size_t SMEM_MAX = 225 * 1024;
        p.bitmap_bytes = 0;
        if (dense) {
            size_t bb = (((size_t)L.dom_size + 31) / 32) * 4;
            bb = (bb + 15) & ~(size_t)15;
            if (bb + 2 * (size_t)p.sub_bytes <= SMEM_MAX) p.bitmap_bytes = (uint32_t)bb;
        }
        uint32_t n_sub = (uint32_t)std::min<size_t>(MT_MAXSUB, (SMEM_MAX - p.bitmap_bytes) / p.sub_bytes);
        if (n_sub < 1) continue;
        const size_t smem_bytes = p.bitmap_bytes + (size_t)n_sub * p.sub_bytes;
        static std::once_flag attr_once;
        static cudaError_t attr_err = cudaSuccess;
        std::call_once(attr_once, [&] { attr_err = cudaFuncSetAttribute((const void*)k_mterms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX); });
        if (attr_err != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "cudaFuncSetAttribute(k_mterms) failed: %s", cudaGetErrorString(attr_err));
        const int threads = (int)n_sub * MT_SUB_THREADS;
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k_mterms, threads, smem_bytes) != cudaSuccess || per_sm < 1)
            return -tagg_fail(TAGG_ERR_CUDA, "k_mterms does not fit an SM (%zu bytes of shared memory)", smem_bytes);

        if (!es.uploads.empty())
            for (uint32_t c = 0; c < es.n_chunks; c++)
                if (cudaStreamWaitEvent(es.st, es.call->chunk_ev[c], 0) != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "stream ordering failed");
        {
            // one persistent launch over the tiles of every segment
            const size_t nseg = es.hsegs.size();
            std::vector<uint32_t> tb(nseg + 1, 0);
            uint64_t tiles = 0;
            for (size_t i = 0; i < nseg; i++) {
                tb[i] = (uint32_t)tiles;
                tiles += ((uint64_t)es.hsegs[i].max_doc + MT_TILE - 1) / MT_TILE;
            }
            if (tiles > 0xffffffffull) continue;
            tb[nseg] = (uint32_t)tiles;
            if (tiles) {
                uint32_t* d_tb = nullptr;
                if (cudaMallocAsync((void**)&d_tb, (nseg + 1) * 4, es.st) != cudaSuccess) return -tagg_fail(TAGG_ERR_OOM, "tile table allocation failed");
                es.temps.push_back(d_tb);
                if (cudaMemcpyAsync(d_tb, es.pin(tb.data(), (nseg + 1) * 4), (nseg + 1) * 4, cudaMemcpyHostToDevice, es.st) != cudaSuccess)
                    return -tagg_fail(TAGG_ERR_CUDA, "tile table upload failed");
                MParams sp = p;
                sp.segs = es.d_segs;
                sp.tile_begin = d_tb;
                sp.n_segs = (uint32_t)nseg;
                sp.n_tiles = (uint32_t)tiles;
                const uint64_t units = (tiles + n_sub - 1) / n_sub;
                const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)es.ctx->sm_count * per_sm, units);
                k_mterms<<<grid, threads, smem_bytes, es.st>>>(sp);
                cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "k_mterms launch failed: %s", cudaGetErrorString(e));
                es.ctx->launches++;
                es.n_launches++;
            }
        }
        es.skip[mem] = 1;
        handled++;
    }
    return handled;
}

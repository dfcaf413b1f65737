// stream.cu — the streaming fast path: K1 (scan_reduce), K2 (terms, dense tables), K3 (histogram).
//
// One persistent launch covers every segment of the call.  The unit of work is a TILE of 2048
// consecutive documents of one segment.  For each tile the bytes of every referenced column
// (256 * num_bits bytes, contiguous and 16-byte aligned in the bit-packed layout) and 256 bytes of
// every bitset (docset, filter_agg docsets, delete bitset) are staged into shared memory by TMA bulk
// copies (cp.async.bulk + mbarrier) through a 3-stage ring, so HBM is read exactly once, fully
// coalesced, while the SM unpacks the previous tile from shared memory.
//
// Per tile each warp owns 128 documents: it folds the bitsets and the value predicates into a 32-bit
// match mask per word (ballot), COMPACTS the matching document indices into a per-warp queue, and
// only then unpacks key / value columns for matched documents — so a 25 %-selective filter costs a
// quarter of the unpack and table instructions (the reference pays a hash probe per matched doc,
// terms.rs:127-132; the doc stream narrowing is filter.rs:100-122 / post_filter.rs:245-249).
//
// Root metrics live in registers and are reduced by warp shuffles; bucket metrics go to dense tables
// in global memory (L2-resident): counts / sums with RED atomics, min / max with a cached
// check-before-atomic (cells only move monotonically, so a stale read can only cause a redundant
// atomic, never a wrong skip).
#include <string.h>

#include <algorithm>

#include "exec.h"

#define ST_THREADS 512
#define ST_WARPS (ST_THREADS / 32)
#define ST_TILE TAGG_TILE_DOCS
#define ST_WORDS_PER_WARP (ST_TILE / 32 / ST_WARPS)  // 4
#define ST_STAGES 3
#define ST_MAXCOLS 6
#define ST_MAXPRED 4
#define ST_MAXBITS (2 + ST_MAXPRED)
#define ST_MAXRG 4
#define ST_MAXBG 3

enum { PR_FILTER = 0, PR_RANGE = 1, PR_LUT = 2, PR_MAIN_RANGE = 3, PR_FILTER_RANGE = 4 };
enum { OPB_SUM = 1, OPB_MIN = 2, OPB_MAX = 4 };
enum { BK_NONE = 0, BK_TERMS = 1, BK_HIST = 2 };

struct SPred {
    int32_t type;
    int32_t scol;     // staged column (RANGE / LUT)
    int32_t filter;   // FILTER: index into DevSegment.filters
    uint32_t pad;
    uint64_t lo, hi;  // RANGE: inclusive code range; LUT: base, number of bits
    const uint8_t* lut;
};
struct SGroup {
    int32_t scol;
    uint32_t kind;
    uint32_t ops;
    uint32_t slot_sum, slot_min, slot_max;
};
struct SParams {
    const DevSegment* segs;
    const uint32_t* tile_prefix;  // n_segs + 1
    const DevPlan* P;
    uint32_t n_segs, n_tiles;
    int32_t n_cols;
    int32_t col_slot[ST_MAXCOLS];   // staged column -> DevSegment.cols index
    uint32_t soff_col[ST_MAXCOLS];  // byte offset inside a stage
    uint32_t soff_bits;             // first bitset slot inside a stage (256 B each): 0 = main, 1 = deleted, 2+i = pred i
    uint32_t stage_bytes;
    int32_t n_preds;
    SPred preds[ST_MAXPRED];
    int32_t n_root_counts;
    uint32_t root_count_slots[2];
    int32_t n_rgroups;
    SGroup rgroups[ST_MAXRG];
    int32_t bucket_mode, key_scol;
    uint64_t dom_min, dom_size;
    double f0, f1;
    uint8_t* present;
    int32_t n_bcounts;
    uint32_t bcount_slots[2];
    int32_t n_bgroups;
    SGroup bgroups[ST_MAXBG];
    int32_t compact;
};

// ---- PTX wrappers: mbarrier + TMA bulk copy ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// value i of a staged column tile (tantivy BitUnpacker::get on shared memory, 32-bit aligned loads)
__device__ __forceinline__ uint64_t sunpack(const uint32_t* __restrict__ s32, uint32_t nb, uint64_t mask, uint32_t i) {
    uint32_t bit = i * nb;
    uint32_t wi = bit >> 5, sh = bit & 31u;
    uint32_t w0 = s32[wi], w1 = s32[wi + 1];
    uint32_t lo = __funnelshift_r(w0, w1, sh);
    uint32_t hi = 0;
    if (nb > 32) {
        uint32_t w2 = s32[wi + 2];
        hi = __funnelshift_r(w1, w2, sh);
    }
    return (((uint64_t)hi << 32) | lo) & mask;
}

struct TileCtx {
    const uint8_t* stage;
    const DevSegment* S;
    uint32_t nb[ST_MAXCOLS];
    uint64_t mask[ST_MAXCOLS];
    uint64_t minv[ST_MAXCOLS];
};

__device__ __forceinline__ uint64_t tile_code(const SParams& p, const TileCtx& t, int scol, uint32_t dl) {
    return sunpack((const uint32_t*)(t.stage + p.soff_col[scol]), t.nb[scol], t.mask[scol], dl) + t.minv[scol];
}

__global__ void __launch_bounds__(ST_THREADS) k_stream(const __grid_constant__ SParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* stages = smem;
    uint16_t* queues = (uint16_t*)(smem + (size_t)ST_STAGES * p.stage_bytes);
    uint64_t* full = (uint64_t*)(queues + ST_WARPS * ST_WORDS_PER_WARP * 32);
    uint32_t* stage_seg = (uint32_t*)(full + ST_STAGES);  // [stage] = segment, [ST_STAGES + stage] = local tile

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < ST_STAGES; s++) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // thread 0 is the producer: locate the tile, arm the stage's barrier with the byte count, issue the copies
    auto issue = [&](uint32_t tile, int stage) {
        uint32_t lo = 0, hi = p.n_segs;  // last seg with prefix <= tile
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (p.tile_prefix[mid] <= tile) lo = mid; else hi = mid;
        }
        const DevSegment* S = p.segs + lo;
        uint32_t lt = tile - p.tile_prefix[lo];
        stage_seg[stage] = lo;
        stage_seg[ST_STAGES + stage] = lt;
        uint8_t* base = stages + (size_t)stage * p.stage_bytes;
        uint32_t bytes = 0;
        for (int c = 0; c < p.n_cols; c++) bytes += (ST_TILE / 8) * S->cols[p.col_slot[c]].num_bits;
        if (S->main.kind == DS_BITSET) bytes += ST_TILE / 8;
        if (S->has_deletes) bytes += ST_TILE / 8;
        for (int i = 0; i < p.n_preds; i++)
            if (p.preds[i].type == PR_FILTER && S->filters[p.preds[i].filter].kind == DS_BITSET) bytes += ST_TILE / 8;
        if (bytes == 0) {  // nothing to stage (e.g. count over AllQuery): complete the phase by hand
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(full + stage)) : "memory");
            return;
        }
        mbar_expect_tx(full + stage, bytes);
        for (int c = 0; c < p.n_cols; c++) {
            const DevColumn& col = S->cols[p.col_slot[c]];
            uint32_t cb = (ST_TILE / 8) * col.num_bits;
            if (cb) tma_bulk_g2s(base + p.soff_col[c], (const uint8_t*)col.words + (size_t)lt * cb, cb, full + stage);
        }
        uint8_t* bits = base + p.soff_bits;
        if (S->main.kind == DS_BITSET) tma_bulk_g2s(bits, (const uint8_t*)S->main.words + (size_t)lt * (ST_TILE / 8), ST_TILE / 8, full + stage);
        if (S->has_deletes) tma_bulk_g2s(bits + 256, (const uint8_t*)S->deleted + (size_t)lt * (ST_TILE / 8), ST_TILE / 8, full + stage);
        for (int i = 0; i < p.n_preds; i++)
            if (p.preds[i].type == PR_FILTER && S->filters[p.preds[i].filter].kind == DS_BITSET)
                tma_bulk_g2s(bits + 512 + 256 * i, (const uint8_t*)S->filters[p.preds[i].filter].words + (size_t)lt * (ST_TILE / 8),
                             ST_TILE / 8, full + stage);
    };

    // per-thread root accumulators
    uint64_t rsum[ST_MAXRG], rmin[ST_MAXRG], rmax[ST_MAXRG];
    bool rseen = false;
#pragma unroll
    for (int g = 0; g < ST_MAXRG; g++) { rsum[g] = 0; rmin[g] = 0; rmax[g] = 0; }
    uint64_t matched = 0;  // lane 0 only

    auto heavy = [&](const TileCtx& t, uint32_t dl) {
        rseen = true;
#pragma unroll
        for (int g = 0; g < ST_MAXRG; g++) {
            if (g < p.n_rgroups) {
                const SGroup& G = p.rgroups[g];
                uint64_t code = tile_code(p, t, G.scol, dl);
                if (G.ops & OPB_SUM) {
                    if (G.kind == TAGG_F64) rsum[g] = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)rsum[g]), code_to_f64(code)));
                    else rsum[g] += code_to_bits(G.kind, code);
                }
                if (G.ops & OPB_MIN) { uint64_t v = ~code; rmin[g] = v > rmin[g] ? v : rmin[g]; }
                if (G.ops & OPB_MAX) rmax[g] = code > rmax[g] ? code : rmax[g];
            }
        }
        if (p.bucket_mode != BK_NONE) {
            uint64_t key = tile_code(p, t, p.key_scol, dl);
            if (p.bucket_mode == BK_HIST) {
                if (!hist_ord(key, p.f0, p.f1, &key)) return;  // NaN or below start: skipped (histogram.rs:138-145)
            }
            uint64_t rel = key - p.dom_min;
            if (key < p.dom_min || rel >= p.dom_size) return;
            if (!p.present[rel]) p.present[rel] = 1;
            for (int c = 0; c < p.n_bcounts; c++) atomicAdd((unsigned long long*)(p.P->slots[p.bcount_slots[c]].acc + rel), 1ull);
#pragma unroll
            for (int g = 0; g < ST_MAXBG; g++) {
                if (g < p.n_bgroups) {
                    const SGroup& G = p.bgroups[g];
                    uint64_t code = tile_code(p, t, G.scol, dl);
                    if (G.ops & OPB_SUM) {
                        uint64_t* a = p.P->slots[G.slot_sum].acc + rel;
                        if (G.kind == TAGG_F64) atomicAdd((double*)a, code_to_f64(code));
                        else atomicAdd((unsigned long long*)a, (unsigned long long)code_to_bits(G.kind, code));
                    }
                    if (G.ops & OPB_MIN) {
                        uint64_t* a = p.P->slots[G.slot_min].acc + rel;
                        uint64_t v = ~code;
                        if (*a < v) atomicMax((unsigned long long*)a, (unsigned long long)v);
                    }
                    if (G.ops & OPB_MAX) {
                        uint64_t* a = p.P->slots[G.slot_max].acc + rel;
                        if (*a < code) atomicMax((unsigned long long*)a, (unsigned long long)code);
                    }
                }
            }
        }
    };

    // prologue: fill ST_STAGES-1 stages
    uint32_t my_first = blockIdx.x, step = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < ST_STAGES - 1; s++) {
            uint64_t t = (uint64_t)my_first + (uint64_t)s * step;
            if (t < p.n_tiles) issue((uint32_t)t, s);
        }
    }
    uint32_t k = 0;
    for (uint64_t tile = my_first; tile < p.n_tiles; tile += step, k++) {
        int stage = k % ST_STAGES;
        uint32_t parity = (k / ST_STAGES) & 1u;
        if (tid == 0) {
            uint64_t nt = tile + (uint64_t)(ST_STAGES - 1) * step;
            if (nt < p.n_tiles) issue((uint32_t)nt, (k + ST_STAGES - 1) % ST_STAGES);
        }
        mbar_wait(full + stage, parity);

        TileCtx t;
        t.stage = stages + (size_t)stage * p.stage_bytes;
        t.S = p.segs + stage_seg[stage];
        const DevSegment& S = *t.S;
        const uint32_t lt = stage_seg[ST_STAGES + stage];
#pragma unroll
        for (int c = 0; c < ST_MAXCOLS; c++) {
            if (c < p.n_cols) {
                const DevColumn& col = S.cols[p.col_slot[c]];
                t.nb[c] = col.num_bits; t.mask[c] = col.mask; t.minv[c] = col.min_value;
            }
        }
        const uint32_t* bits = (const uint32_t*)(t.stage + p.soff_bits);
        uint16_t* q = queues + warp * (ST_WORDS_PER_WARP * 32);
        uint32_t nq = 0;
#pragma unroll
        for (int j = 0; j < ST_WORDS_PER_WARP; j++) {
            uint32_t wi = warp * ST_WORDS_PER_WARP + j;
            uint32_t dl = wi * 32 + lane;                 // doc index inside the tile
            uint64_t base_doc = (uint64_t)lt * ST_TILE + (uint64_t)wi * 32;
            uint32_t m;
            if (base_doc + 32 <= S.max_doc) m = 0xffffffffu;
            else if (base_doc >= S.max_doc) m = 0;
            else m = (1u << (uint32_t)(S.max_doc - base_doc)) - 1u;
            if (S.main.kind == DS_BITSET) m &= bits[wi];
            if (S.has_deletes) m &= ~bits[64 + wi];   // searcher.rs:41-46
            for (int i = 0; i < p.n_preds && m; i++) {
                const SPred& pr = p.preds[i];
                if (pr.type == PR_FILTER) {
                    const DevDocset& fd = S.filters[pr.filter];
                    if (fd.kind == DS_BITSET) m &= bits[128 + 64 * i + wi];
                    else if (fd.kind != DS_ALL) m = 0;
                } else {
                    uint64_t lo = pr.lo, hi = pr.hi;
                    if (pr.type == PR_MAIN_RANGE) { lo = S.main.lo; hi = S.main.hi; }
                    else if (pr.type == PR_FILTER_RANGE) { lo = S.filters[pr.filter].lo; hi = S.filters[pr.filter].hi; }
                    uint64_t code = tile_code(p, t, pr.scol, dl);
                    bool ok;
                    if (pr.type == PR_LUT) {
                        uint64_t r = code - lo;
                        ok = code >= lo && r < hi && ((pr.lut[r >> 3] >> (r & 7)) & 1);
                    } else {
                        ok = code >= lo && code <= hi;
                    }
                    m &= __ballot_sync(0xffffffffu, ok);
                }
            }
            if (lane == 0) matched += __popc(m);
            if (p.compact) {
                if ((m >> lane) & 1u) q[nq + __popc(m & ((1u << lane) - 1u))] = (uint16_t)dl;
                nq += __popc(m);
            } else if ((m >> lane) & 1u) {
                heavy(t, dl);
            }
        }
        if (p.compact) {
            __syncwarp();
            for (uint32_t jj = lane; jj < nq; jj += 32) heavy(t, q[jj]);
        }
        __syncthreads();  // every warp is done with this stage: the producer may refill it
    }

    // fold the root accumulators (warp shuffle, then one atomic per warp)
    for (int c = 0; c < p.n_root_counts; c++)
        if (lane == 0 && matched) atomicAdd((unsigned long long*)p.P->slots[p.root_count_slots[c]].acc, (unsigned long long)matched);
    uint32_t any = __ballot_sync(0xffffffffu, rseen);
#pragma unroll
    for (int g = 0; g < ST_MAXRG; g++) {
        if (g < p.n_rgroups) {
            const SGroup& G = p.rgroups[g];
            uint64_t s = rsum[g], mn = rmin[g], mx = rmax[g];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                uint64_t s2 = __shfl_xor_sync(0xffffffffu, s, o), mn2 = __shfl_xor_sync(0xffffffffu, mn, o), mx2 = __shfl_xor_sync(0xffffffffu, mx, o);
                if (G.kind == TAGG_F64) s = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)s), __longlong_as_double((long long)s2)));
                else s += s2;
                mn = mn2 > mn ? mn2 : mn;
                mx = mx2 > mx ? mx2 : mx;
            }
            if (lane == 0 && any) {
                if (G.ops & OPB_SUM) {
                    const DevSlot& sl = p.P->slots[G.slot_sum];
                    if (G.kind == TAGG_F64) atomicAdd((double*)sl.acc, __longlong_as_double((long long)s));
                    else atomicAdd((unsigned long long*)sl.acc, (unsigned long long)s);
                    sl.seen[0] = 1;
                }
                if (G.ops & OPB_MIN) { const DevSlot& sl = p.P->slots[G.slot_min]; atomicMax((unsigned long long*)sl.acc, (unsigned long long)mn); sl.seen[0] = 1; }
                if (G.ops & OPB_MAX) { const DevSlot& sl = p.P->slots[G.slot_max]; atomicMax((unsigned long long*)sl.acc, (unsigned long long)mx); sl.seen[0] = 1; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host: does the plan have the flat streaming shape?  [filters / post-filters]* -> (root metrics..., one
// dense TERMS | HISTOGRAM over leaf metrics), every column single-valued.
// ------------------------------------------------------------------------------------------------------
struct Shape {
    SParams sp;
    std::vector<int> staged;  // DevSegment.cols slot per staged column
    int stage_col(int slot) {
        for (size_t i = 0; i < staged.size(); i++)
            if (staged[i] == slot) return (int)i;
        if (staged.size() >= ST_MAXCOLS) return -1;
        staged.push_back(slot);
        return (int)staged.size() - 1;
    }
};

static bool add_fold(Shape& sh, SGroup* groups, int32_t& n, int maxn, const PlanMeta& m, int node) {
    const tagg_node& nd = m.nodes[node];
    if (nd.multi) return false;
    int scol = sh.stage_col(m.col_slot[node]);
    if (scol < 0) return false;
    uint32_t bit = nd.op == TAGG_OP_SUM ? OPB_SUM : nd.op == TAGG_OP_MIN ? OPB_MIN : OPB_MAX;
    SGroup* G = nullptr;
    for (int i = 0; i < n; i++)
        if (groups[i].scol == scol && groups[i].kind == nd.kind) G = &groups[i];
    if (!G) {
        if (n >= maxn) return false;
        G = &groups[n++];
        memset(G, 0, sizeof(*G));
        G->scol = scol;
        G->kind = nd.kind;
    }
    if (G->ops & bit) return false;  // the same op twice on one column: leave it to the generic kernel
    G->ops |= bit;
    uint32_t slot = (uint32_t)m.slot_of[node];
    if (bit == OPB_SUM) G->slot_sum = slot; else if (bit == OPB_MIN) G->slot_min = slot; else G->slot_max = slot;
    return true;
}

int stream_try(ExecState& es) {
    const PlanMeta& m = *es.meta;
    if (es.segs.empty() || !m.pct_node.empty()) return 0;
    Shape sh;
    SParams& sp = sh.sp;
    memset(&sp, 0, sizeof(sp));
    sp.key_scol = -1;
    uint32_t n_nodes = (uint32_t)m.nodes.size();

    // docsets must have one kind per position across segments (ALL / BITSET mixes are handled per segment)
    const DevSegment& S0 = es.hsegs[0];
    for (auto& hs : es.hsegs) {
        if (hs.main.kind == DS_IDS) return 0;  // sparse id lists: the gather (generic) kernel is the right tool
        if ((hs.main.kind == DS_RANGE) != (S0.main.kind == DS_RANGE)) return 0;
        if (hs.main.kind == DS_RANGE && hs.main.col != S0.main.col) return 0;
        for (uint32_t f = 0; f < m.n_filters; f++) {
            if ((hs.filters[f].kind == DS_RANGE) != (S0.filters[f].kind == DS_RANGE)) return 0;
            if (hs.filters[f].kind == DS_RANGE && hs.filters[f].col != S0.filters[f].col) return 0;
        }
    }
    if (S0.main.kind == DS_RANGE) {
        SPred& pr = sp.preds[sp.n_preds++];
        pr.type = PR_MAIN_RANGE;
        pr.scol = sh.stage_col(S0.main.col);
    }
    uint32_t node = 0;
    while (node < n_nodes && (m.nodes[node].op == TAGG_OP_FILTER || m.nodes[node].op == TAGG_OP_POST_FILTER)) {
        const tagg_node& nd = m.nodes[node];
        if (sp.n_preds >= ST_MAXPRED) return 0;
        SPred& pr = sp.preds[sp.n_preds];
        if (nd.op == TAGG_OP_FILTER) {
            pr.filter = (int32_t)nd.aux;
            if (S0.filters[nd.aux].kind == DS_RANGE) {
                pr.type = PR_FILTER_RANGE;
                pr.scol = sh.stage_col(S0.filters[nd.aux].col);
                if (pr.scol < 0) return 0;
            } else {
                pr.type = PR_FILTER;
            }
        } else {
            if (nd.multi) return 0;
            pr.type = nd.pred == TAGG_PRED_LUT ? PR_LUT : PR_RANGE;
            pr.scol = sh.stage_col(m.col_slot[node]);
            if (pr.scol < 0) return 0;
            pr.lo = nd.u0;
            pr.hi = nd.u1;
            pr.lut = nd.pred == TAGG_PRED_LUT ? es.plan->d_blobs[nd.aux] : nullptr;
        }
        sp.n_preds++;
        node++;
    }
    if (node >= n_nodes) return 0;
    // FILTER nodes deeper in the tree are not part of the flat shape
    for (uint32_t i = node; i < n_nodes; i++)
        if (m.nodes[i].op == TAGG_OP_FILTER || m.nodes[i].op == TAGG_OP_POST_FILTER) return 0;

    std::vector<uint32_t> members;
    if (m.nodes[node].op == TAGG_OP_TUPLE) {
        for (uint32_t c = node + 1; c < m.end[node]; c = m.end[c]) members.push_back(c);
    } else {
        members.push_back(node);
    }
    for (uint32_t mem : members) {
        const tagg_node& nd = m.nodes[mem];
        if (nd.op == TAGG_OP_COUNT) {
            if (sp.n_root_counts >= 2) return 0;
            sp.root_count_slots[sp.n_root_counts++] = (uint32_t)m.slot_of[mem];
        } else if (nd.op == TAGG_OP_SUM || nd.op == TAGG_OP_MIN || nd.op == TAGG_OP_MAX) {
            if (!add_fold(sh, sp.rgroups, sp.n_rgroups, ST_MAXRG, m, (int)mem)) return 0;
        } else if (nd.op == TAGG_OP_TERMS || nd.op == TAGG_OP_HISTOGRAM) {
            if (sp.bucket_mode != BK_NONE || nd.multi) return 0;
            int sc = m.own_scope[mem];
            const ScopeLayout& L = es.scopes[sc];
            if (L.mode != SCOPE_DENSE) return 0;
            sp.bucket_mode = nd.op == TAGG_OP_TERMS ? BK_TERMS : BK_HIST;
            sp.key_scol = sh.stage_col(m.col_slot[mem]);
            if (sp.key_scol < 0) return 0;
            sp.dom_min = L.dom_min;
            sp.dom_size = L.dom_size;
            sp.f0 = nd.f0;
            sp.f1 = nd.f1;
            sp.present = es.arena + L.off_present;
            uint32_t sub = mem + 1;
            std::vector<uint32_t> leaves;
            if (m.nodes[sub].op == TAGG_OP_TUPLE) {
                for (uint32_t c = sub + 1; c < m.end[sub]; c = m.end[c]) leaves.push_back(c);
            } else {
                leaves.push_back(sub);
            }
            for (uint32_t lf : leaves) {
                const tagg_node& ln = m.nodes[lf];
                if (ln.op == TAGG_OP_COUNT) {
                    if (sp.n_bcounts >= 2) return 0;
                    sp.bcount_slots[sp.n_bcounts++] = (uint32_t)m.slot_of[lf];
                } else if (ln.op == TAGG_OP_SUM || ln.op == TAGG_OP_MIN || ln.op == TAGG_OP_MAX) {
                    if (!add_fold(sh, sp.bgroups, sp.n_bgroups, ST_MAXBG, m, (int)lf)) return 0;
                } else {
                    return 0;
                }
            }
        } else {
            return 0;
        }
    }

    // stage layout: every staged column sized for its widest segment
    sp.n_cols = (int32_t)sh.staged.size();
    uint32_t off = 0;
    for (int c = 0; c < sp.n_cols; c++) {
        sp.col_slot[c] = sh.staged[c];
        uint32_t maxnb = 0;
        for (auto& hs : es.hsegs) maxnb = std::max(maxnb, hs.cols[sh.staged[c]].num_bits);
        sp.soff_col[c] = off;
        off += (ST_TILE / 8) * maxnb + 16;
    }
    off = (off + 127) & ~127u;
    sp.soff_bits = off;
    off += 256 * ST_MAXBITS;
    sp.stage_bytes = (off + 127) & ~127u;
    size_t smem_bytes = (size_t)ST_STAGES * sp.stage_bytes + ST_WARPS * ST_WORDS_PER_WARP * 32 * 2 + ST_STAGES * 8 + 2 * ST_STAGES * 4 + 64;
    if (smem_bytes > 200 * 1024) return 0;

    // tile table
    std::vector<uint32_t> prefix(es.hsegs.size() + 1, 0);
    for (size_t i = 0; i < es.hsegs.size(); i++) {
        uint64_t tiles = ((uint64_t)es.hsegs[i].max_doc + ST_TILE - 1) / ST_TILE;
        uint64_t nx = prefix[i] + tiles;
        if (nx > 0xffffffffull) return 0;
        prefix[i + 1] = (uint32_t)nx;
    }
    sp.n_segs = (uint32_t)es.hsegs.size();
    sp.n_tiles = prefix.back();
    // the bucket slots' Option flags coincide with bucket existence in the flat shape: alias them
    if (sp.bucket_mode != BK_NONE) {
        int sc = -1;
        for (size_t s = 1; s < es.scopes.size(); s++) sc = (int)s;
        for (size_t k = 0; k < es.slots.size(); k++)
            if (m.scope_of[m.slot_node[k]] == sc) es.slots[k].off_seen = es.scopes[sc].off_present;
    }
    es.path_used = 2;
    if (sp.n_tiles == 0) return 1;
    uint32_t* d_prefix = nullptr;
    if (cudaMallocAsync((void**)&d_prefix, prefix.size() * 4, es.st) != cudaSuccess) return -tagg_fail(TAGG_ERR_OOM, "tile table allocation failed");
    es.temps.push_back(d_prefix);
    if (cudaMemcpyAsync(d_prefix, prefix.data(), prefix.size() * 4, cudaMemcpyHostToDevice, es.st) != cudaSuccess)
        return -tagg_fail(TAGG_ERR_CUDA, "tile table upload failed");
    sp.tile_prefix = d_prefix;
    sp.segs = es.d_segs;
    sp.P = es.d_plan;
    // compaction pays when most documents are filtered out; with no narrowing it is pure overhead
    bool narrowing = sp.n_preds > 0;
    for (auto& hs : es.hsegs) narrowing = narrowing || hs.main.kind == DS_BITSET || hs.has_deletes;
    sp.compact = narrowing ? 1 : 0;

    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return -tagg_fail(TAGG_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
        attr_set = true;
    }
    int per_sm = (int)std::min<size_t>(2048 / ST_THREADS, (227 * 1024) / (smem_bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)es.ctx->sm_count * per_sm, sp.n_tiles);
    k_stream<<<grid, ST_THREADS, smem_bytes, es.st>>>(sp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "k_stream launch failed: %s", cudaGetErrorString(e));
    es.ctx->launches++;
    es.n_launches++;
    return 1;
}

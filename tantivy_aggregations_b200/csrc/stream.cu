// stream.cu — the streaming fast path: K1 (scan_reduce), K2 (terms, dense tables), K3 (histogram).
//
// One persistent launch covers every segment of the call.  The unit of work is a TILE of 2048
// consecutive documents of one segment.  For each tile the bytes of every referenced column
// (256 * num_bits bytes, contiguous and 16-byte aligned in the bit-packed layout) and 256 bytes of
// every bitset (docset, filter_agg docsets, delete bitset) are staged into shared memory by TMA bulk
// copies (cp.async.bulk + mbarrier) through a 3-stage ring, so HBM is read exactly once, fully
// coalesced, while the SM unpacks the previous tile from shared memory.
//
// Per tile each warp owns 256 documents = 8 bitset words:
//   phase 1  lanes 0..7 AND the staged bitset words (docset, ~deletes, filter docsets) of "their"
//            word into a match mask; value predicates (post_filter / COLUMN_RANGE) are evaluated per
//            document and folded in with ballots;
//   phase 2  the set bits are COMPACTED into a per-warp queue of document indices (warp scan of the
//            popcounts + one predicated shared store per word);
//   phase 3  full warps drain the queue: only matched documents pay the unpack of key / value
//            columns and the table updates (the reference pays a hash probe per matched doc,
//            terms.rs:127-132; the doc-stream narrowing is filter.rs:100-122 / post_filter.rs:245-249).
// With nothing narrowing the doc stream (AllQuery, no deletes) phase 2 is skipped.
//
// Root metrics live in registers and are reduced by warp shuffles; bucket metrics go to dense tables
// in global memory (L2-resident): counts / sums with RED atomics, min / max with a cached
// check-before-atomic (cells only move monotonically, so a stale read can only cause a redundant
// atomic, never a wrong skip).  The kernel is a template over the number of root / bucket column
// groups so every per-tile column descriptor lives in registers.
#include <string.h>

#include <algorithm>

#include "exec.h"

#define ST_THREADS 256
#define ST_WARPS (ST_THREADS / 32)
#define ST_TILE TAGG_TILE_DOCS
#define ST_WORDS_PER_WARP (ST_TILE / 32 / ST_WARPS)  // 8
#define ST_DOCS_PER_WARP (ST_WORDS_PER_WARP * 32)    // 256
#define ST_STAGES 3
#define ST_MAXCOLS 6
#define ST_MAXPRED 4
#define ST_MAXBITS (2 + ST_MAXPRED)
#define ST_MAXRG 4
#define ST_MAXBG 3

enum { PR_FILTER = 0, PR_RANGE = 1, PR_LUT = 2 };
enum { OPB_SUM = 1, OPB_MIN = 2, OPB_MAX = 4 };
enum { BK_NONE = 0, BK_TERMS = 1, BK_HIST = 2 };
enum { SF_MAIN_BITS = 1, SF_DELETES = 2, SF_PRED_BITS0 = 4 /* << i */, SF_PRED_NONE0 = 256 /* << i */ };

// Everything the kernel needs to know about one segment, prepared on the host.
struct SegDesc {
    uint32_t tile_begin, max_doc, flags, pad;
    const uint8_t* col_ptr[ST_MAXCOLS];
    uint64_t minv[ST_MAXCOLS];
    uint32_t nb[ST_MAXCOLS];
    const uint8_t* bits_ptr[ST_MAXBITS];  // 0 = main docset, 1 = deleted, 2+i = filter docset of pred i
    uint64_t pred_lo[ST_MAXPRED], pred_hi[ST_MAXPRED];
};
struct SGroup {
    int32_t scol;
    uint32_t kind;
    uint32_t ops;
    uint32_t pad;
    uint64_t *acc_sum, *acc_min, *acc_max;
    uint8_t *seen_sum, *seen_min, *seen_max;
};
struct SParams {
    const SegDesc* segs;
    uint32_t n_segs, n_tiles;
    int32_t n_cols;
    uint32_t soff_col[ST_MAXCOLS];  // byte offset of each staged column inside a stage
    uint32_t soff_bits;             // bitset slots (256 B each) inside a stage
    uint32_t stage_bytes;
    int32_t n_preds;
    int32_t pred_type[ST_MAXPRED];
    int32_t pred_scol[ST_MAXPRED];
    const uint8_t* pred_lut[ST_MAXPRED];
    int32_t n_root_counts;
    uint64_t* root_count_acc[2];
    SGroup rgroups[ST_MAXRG];
    int32_t key_scol;
    uint64_t dom_min, dom_size;
    double f0, f1;
    uint8_t* present;  // nullptr: derived from the bucket counts after the kernel
    int32_t n_bcounts;
    uint64_t* bcount_acc[2];
    SGroup bgroups[ST_MAXBG];
};

// ---- PTX wrappers: mbarrier + TMA bulk copy ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// A staged column of the current tile, held in registers.
struct TCol {
    const uint32_t* s32;
    uint64_t minv, mask;
    uint32_t nb;
};
__device__ __forceinline__ TCol tcol(const SParams& p, const SegDesc* S, const uint8_t* stage, int scol) {
    TCol c;
    c.s32 = (const uint32_t*)(stage + p.soff_col[scol]);
    c.nb = S->nb[scol];
    c.minv = S->minv[scol];
    c.mask = c.nb == 64 ? ~0ull : ((1ull << c.nb) - 1ull);
    return c;
}
// value i of a staged column tile (tantivy BitUnpacker::get on shared memory, 32-bit aligned loads) -> code
__device__ __forceinline__ uint64_t tget(const TCol& c, uint32_t i) {
    uint32_t bit = i * c.nb;
    uint32_t wi = bit >> 5, sh = bit & 31u;
    uint32_t w0 = c.s32[wi], w1 = c.s32[wi + 1];
    uint32_t lo = __funnelshift_r(w0, w1, sh);
    uint32_t hi = 0;
    if (c.nb > 32) {
        uint32_t w2 = c.s32[wi + 2];
        hi = __funnelshift_r(w1, w2, sh);
    }
    return ((((uint64_t)hi << 32) | lo) & c.mask) + c.minv;
}

template <int BUCKET, int NBG, int NRG, bool COMPACT>
__global__ void __launch_bounds__(ST_THREADS) k_stream(const __grid_constant__ SParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* stages = smem;
    uint16_t* queues = (uint16_t*)(smem + (size_t)ST_STAGES * p.stage_bytes);
    uint64_t* full = (uint64_t*)(queues + ST_WARPS * ST_DOCS_PER_WARP);
    uint32_t* stage_seg = (uint32_t*)(full + ST_STAGES);  // [stage] = segment, [ST_STAGES + stage] = local tile

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    if (tid == 0) {
        for (int s = 0; s < ST_STAGES; s++) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // thread 0 is the producer: tiles are visited in ascending order, so the segment cursor only advances
    uint32_t cur_seg = 0;
    auto issue = [&](uint32_t tile, int stage) {
        while (cur_seg + 1 < p.n_segs && p.segs[cur_seg + 1].tile_begin <= tile) cur_seg++;
        const SegDesc* S = p.segs + cur_seg;
        uint32_t lt = tile - S->tile_begin;
        stage_seg[stage] = cur_seg;
        stage_seg[ST_STAGES + stage] = lt;
        uint8_t* base = stages + (size_t)stage * p.stage_bytes;
        uint32_t flags = S->flags;
        uint32_t bytes = 0;
        for (int c = 0; c < p.n_cols; c++) bytes += (ST_TILE / 8) * S->nb[c];
        bytes += (ST_TILE / 8) * __popc(flags & 0xffu);
        if (bytes == 0) {  // nothing to stage (count over AllQuery): complete the phase by hand
            mbar_arrive(full + stage);
            return;
        }
        mbar_expect_tx(full + stage, bytes);
        for (int c = 0; c < p.n_cols; c++) {
            uint32_t cb = (ST_TILE / 8) * S->nb[c];
            if (cb) tma_bulk_g2s(base + p.soff_col[c], S->col_ptr[c] + (size_t)lt * cb, cb, full + stage);
        }
        uint8_t* bits = base + p.soff_bits;
        for (int b = 0; b < ST_MAXBITS; b++)
            if (flags & (1u << b)) tma_bulk_g2s(bits + 256 * b, S->bits_ptr[b] + (size_t)lt * (ST_TILE / 8), ST_TILE / 8, full + stage);
    };

    // per-thread root accumulators
    uint64_t rsum[NRG ? NRG : 1], rmin[NRG ? NRG : 1], rmax[NRG ? NRG : 1];
    bool rseen = false;
#pragma unroll
    for (int g = 0; g < NRG; g++) { rsum[g] = 0; rmin[g] = 0; rmax[g] = 0; }
    uint32_t matched = 0;  // every lane holds the warp's count

    uint32_t my_first = blockIdx.x, step = gridDim.x;
    if (tid == 0) {
        for (int s = 0; s < ST_STAGES - 1; s++) {
            uint64_t t = (uint64_t)my_first + (uint64_t)s * step;
            if (t < p.n_tiles) issue((uint32_t)t, s);
        }
    }
    uint32_t k = 0;
    for (uint64_t tile = my_first; tile < p.n_tiles; tile += step, k++) {
        const int stage = k % ST_STAGES;
        const uint32_t parity = (k / ST_STAGES) & 1u;
        if (tid == 0) {
            uint64_t nt = tile + (uint64_t)(ST_STAGES - 1) * step;
            if (nt < p.n_tiles) issue((uint32_t)nt, (k + ST_STAGES - 1) % ST_STAGES);
        }
        mbar_wait(full + stage, parity);

        const uint8_t* sbase = stages + (size_t)stage * p.stage_bytes;
        const SegDesc* S = p.segs + stage_seg[stage];
        const uint32_t lt = stage_seg[ST_STAGES + stage];
        const uint32_t flags = S->flags;
        const uint32_t* bits = (const uint32_t*)(sbase + p.soff_bits);
        // documents of this tile that exist
        const uint64_t tile_doc0 = (uint64_t)lt * ST_TILE;
        const uint32_t n_valid = (uint32_t)min((uint64_t)ST_TILE, (uint64_t)S->max_doc - tile_doc0);

        // ---- phase 1: one match-mask word per lane (lanes 0..7) ------------------------------------
        uint32_t m = 0;
        if (lane < ST_WORDS_PER_WARP) {
            uint32_t wi = warp * ST_WORDS_PER_WARP + lane;
            uint32_t d0 = wi * 32;
            m = d0 + 32 <= n_valid ? 0xffffffffu : (d0 >= n_valid ? 0u : ((1u << (n_valid - d0)) - 1u));
            if (flags & SF_MAIN_BITS) m &= bits[wi];
            if (flags & SF_DELETES) m &= ~bits[64 + wi];  // searcher.rs:41-46
            for (int i = 0; i < p.n_preds; i++) {
                if (flags & (SF_PRED_BITS0 << i)) m &= bits[128 + 64 * i + wi];
                if (flags & (SF_PRED_NONE0 << i)) m = 0;
            }
        }
        // value predicates: evaluated per document, folded in with ballots
        for (int i = 0; i < p.n_preds; i++) {
            const int type = p.pred_type[i];
            if (type == PR_FILTER) continue;
            const TCol pc = tcol(p, S, sbase, p.pred_scol[i]);
            const uint64_t lo = S->pred_lo[i], hi = S->pred_hi[i];
            const uint8_t* lut = p.pred_lut[i];
#pragma unroll
            for (int j = 0; j < ST_WORDS_PER_WARP; j++) {
                uint32_t mj = __shfl_sync(0xffffffffu, m, j);
                if (mj) {
                    uint64_t code = tget(pc, warp * ST_DOCS_PER_WARP + j * 32 + lane);
                    bool ok;
                    if (type == PR_LUT) {
                        uint64_t r = code - lo;
                        ok = code >= lo && r < hi && ((lut[r >> 3] >> (r & 7)) & 1);
                    } else {
                        ok = code >= lo && code <= hi;
                    }
                    mj &= __ballot_sync(0xffffffffu, ok);
                    if (lane == j) m = mj;
                }
            }
        }

        // per-tile column descriptors of the roles this instantiation has, in registers
        TCol kc, bc[NBG ? NBG : 1], rc[NRG ? NRG : 1];
        if (BUCKET != BK_NONE) kc = tcol(p, S, sbase, p.key_scol);
#pragma unroll
        for (int g = 0; g < NBG; g++) bc[g] = tcol(p, S, sbase, p.bgroups[g].scol);
#pragma unroll
        for (int g = 0; g < NRG; g++) rc[g] = tcol(p, S, sbase, p.rgroups[g].scol);

        auto heavy = [&](uint32_t dl) {
            rseen = true;
#pragma unroll
            for (int g = 0; g < NRG; g++) {
                const SGroup& G = p.rgroups[g];
                if (G.ops) {
                    uint64_t code = tget(rc[g], dl);
                    if (G.ops & OPB_SUM) {
                        if (G.kind == TAGG_F64) rsum[g] = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)rsum[g]), code_to_f64(code)));
                        else rsum[g] += code_to_bits(G.kind, code);
                    }
                    if (G.ops & OPB_MIN) { uint64_t v = ~code; rmin[g] = v > rmin[g] ? v : rmin[g]; }
                    if (G.ops & OPB_MAX) rmax[g] = code > rmax[g] ? code : rmax[g];
                }
            }
            if (BUCKET != BK_NONE) {
                uint64_t key = tget(kc, dl);
                if (BUCKET == BK_HIST) {
                    if (!hist_ord(key, p.f0, p.f1, &key)) return;  // NaN or below start: skipped (histogram.rs:138-145)
                }
                uint64_t rel = key - p.dom_min;
                if (key < p.dom_min || rel >= p.dom_size) return;
                if (p.present && !p.present[rel]) p.present[rel] = 1;
                if (p.n_bcounts > 0) atomicAdd((unsigned long long*)(p.bcount_acc[0] + rel), 1ull);
                if (p.n_bcounts > 1) atomicAdd((unsigned long long*)(p.bcount_acc[1] + rel), 1ull);
#pragma unroll
                for (int g = 0; g < NBG; g++) {
                    const SGroup& G = p.bgroups[g];
                    uint64_t code = tget(bc[g], dl);
                    if (G.ops & OPB_SUM) {
                        if (G.kind == TAGG_F64) atomicAdd((double*)(G.acc_sum + rel), code_to_f64(code));
                        else atomicAdd((unsigned long long*)(G.acc_sum + rel), (unsigned long long)code_to_bits(G.kind, code));
                    }
                    if (G.ops & OPB_MIN) {
                        uint64_t v = ~code;
                        if (G.acc_min[rel] < v) atomicMax((unsigned long long*)(G.acc_min + rel), (unsigned long long)v);
                    }
                    if (G.ops & OPB_MAX) {
                        if (G.acc_max[rel] < code) atomicMax((unsigned long long*)(G.acc_max + rel), (unsigned long long)code);
                    }
                }
            }
        };

        if (COMPACT) {
            // ---- phase 2: compact the set bits of the 8 words into the warp's queue ------------------
            uint32_t incl = __popc(m);
#pragma unroll
            for (int o = 1; o < ST_WORDS_PER_WARP; o <<= 1) {
                uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += up;
            }
            const uint32_t excl = incl - __popc(m);
            const uint32_t nq = __shfl_sync(0xffffffffu, incl, ST_WORDS_PER_WARP - 1);
            uint16_t* q = queues + warp * ST_DOCS_PER_WARP;
#pragma unroll
            for (int j = 0; j < ST_WORDS_PER_WARP; j++) {
                uint32_t mj = __shfl_sync(0xffffffffu, m, j);
                uint32_t oj = __shfl_sync(0xffffffffu, excl, j);
                if ((mj >> lane) & 1u) q[oj + __popc(mj & lt_mask)] = (uint16_t)(j * 32 + lane);
            }
            matched += nq;
            __syncwarp();
            // ---- phase 3: full warps drain the queue --------------------------------------------------
            for (uint32_t jj = lane; jj < nq; jj += 32) heavy(warp * ST_DOCS_PER_WARP + q[jj]);
        } else {
#pragma unroll
            for (int j = 0; j < ST_WORDS_PER_WARP; j++) {
                uint32_t mj = __shfl_sync(0xffffffffu, m, j);
                matched += __popc(mj);
                if ((mj >> lane) & 1u) heavy(warp * ST_DOCS_PER_WARP + j * 32 + lane);
            }
        }
        __syncthreads();  // every warp is done with this stage: the producer may refill it
    }

    // fold the root accumulators (warp shuffle, then one atomic per warp)
    if (lane == 0 && matched) {
        if (p.n_root_counts > 0) atomicAdd((unsigned long long*)p.root_count_acc[0], (unsigned long long)matched);
        if (p.n_root_counts > 1) atomicAdd((unsigned long long*)p.root_count_acc[1], (unsigned long long)matched);
    }
    if (NRG > 0) {
        uint32_t any = __ballot_sync(0xffffffffu, rseen);
#pragma unroll
        for (int g = 0; g < NRG; g++) {
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
                    if (G.kind == TAGG_F64) atomicAdd((double*)G.acc_sum, __longlong_as_double((long long)s));
                    else atomicAdd((unsigned long long*)G.acc_sum, (unsigned long long)s);
                    *G.seen_sum = 1;
                }
                if (G.ops & OPB_MIN) { atomicMax((unsigned long long*)G.acc_min, (unsigned long long)mn); *G.seen_min = 1; }
                if (G.ops & OPB_MAX) { atomicMax((unsigned long long*)G.acc_max, (unsigned long long)mx); *G.seen_max = 1; }
            }
        }
    }
}

// bucket existence from the bucket counts (when the plan has a count under the bucket node the kernel
// does not maintain `present` itself: a bucket exists iff its count is non-zero)
__global__ void k_present_from_counts(const uint64_t* __restrict__ counts, uint8_t* __restrict__ present, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        present[i] = counts[i] != 0;
}

// ------------------------------------------------------------------------------------------------------
// host: does the plan have the flat streaming shape?  [filters / post-filters]* -> (root metrics..., one
// dense TERMS | HISTOGRAM over leaf metrics), every column single-valued.
// ------------------------------------------------------------------------------------------------------
typedef void (*stream_fn)(const SParams);
template <int BUCKET, int NBG, int NRG>
static stream_fn pick_compact(bool compact) {
    return compact ? (stream_fn)k_stream<BUCKET, NBG, NRG, true> : (stream_fn)k_stream<BUCKET, NBG, NRG, false>;
}
template <int BUCKET, int NBG>
static stream_fn pick_nrg(int nrg, bool compact) {
    switch (nrg) {
        case 0: return pick_compact<BUCKET, NBG, 0>(compact);
        case 1: return pick_compact<BUCKET, NBG, 1>(compact);
        default: return pick_compact<BUCKET, NBG, ST_MAXRG>(compact);
    }
}
template <int BUCKET>
static stream_fn pick_nbg(int nbg, int nrg, bool compact) {
    switch (nbg) {
        case 0: return pick_nrg<BUCKET, 0>(nrg, compact);
        case 1: return pick_nrg<BUCKET, 1>(nrg, compact);
        case 2: return pick_nrg<BUCKET, 2>(nrg, compact);
        default: return pick_nrg<BUCKET, 3>(nrg, compact);
    }
}
static stream_fn pick_kernel(int bucket, int nbg, int nrg, bool compact) {
    switch (bucket) {
        case BK_NONE: return pick_nrg<BK_NONE, 0>(nrg, compact);
        case BK_TERMS: return pick_nbg<BK_TERMS>(nbg, nrg, compact);
        default: return pick_nbg<BK_HIST>(nbg, nrg, compact);
    }
}

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

static bool add_fold(ExecState& es, Shape& sh, SGroup* groups, int& n, int maxn, int node) {
    const PlanMeta& m = *es.meta;
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
    const SlotLayout& SL = es.slots[m.slot_of[node]];
    uint64_t* acc = (uint64_t*)(es.arena + SL.off_acc);
    uint8_t* seen = es.arena + SL.off_seen;
    if (bit == OPB_SUM) { G->acc_sum = acc; G->seen_sum = seen; }
    else if (bit == OPB_MIN) { G->acc_min = acc; G->seen_min = seen; }
    else { G->acc_max = acc; G->seen_max = seen; }
    return true;
}

int stream_try(ExecState& es) {
    const PlanMeta& m = *es.meta;
    if (es.segs.empty() || !m.pct_node.empty()) return 0;
    Shape sh;
    SParams& sp = sh.sp;
    memset(&sp, 0, sizeof(sp));
    sp.key_scol = -1;
    const uint32_t n_nodes = (uint32_t)m.nodes.size();
    const size_t nseg = es.hsegs.size();

    // docset kinds must agree across segments where a column is involved (ALL / BITSET mixes are per segment)
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
    // predicate list: pred_src[i] = -2 main docset range, -1 plan node constant, f >= 0 filter docset f
    std::vector<int> pred_src;
    auto add_pred = [&](int type, int scol, int src, uint64_t lo, uint64_t hi, const uint8_t* lut) -> bool {
        if (sp.n_preds >= ST_MAXPRED) return false;
        int i = sp.n_preds++;
        sp.pred_type[i] = type;
        sp.pred_scol[i] = scol;
        sp.pred_lut[i] = lut;
        pred_src.push_back(src);
        (void)lo; (void)hi;
        return true;
    };
    std::vector<std::pair<uint64_t, uint64_t>> pred_const(ST_MAXPRED);
    if (S0.main.kind == DS_RANGE) {
        int sc = sh.stage_col(S0.main.col);
        if (sc < 0 || !add_pred(PR_RANGE, sc, -2, 0, 0, nullptr)) return 0;
    }
    uint32_t node = 0;
    while (node < n_nodes && (m.nodes[node].op == TAGG_OP_FILTER || m.nodes[node].op == TAGG_OP_POST_FILTER)) {
        const tagg_node& nd = m.nodes[node];
        if (nd.op == TAGG_OP_FILTER) {
            if (S0.filters[nd.aux].kind == DS_RANGE) {
                int sc = sh.stage_col(S0.filters[nd.aux].col);
                if (sc < 0 || !add_pred(PR_RANGE, sc, (int)nd.aux, 0, 0, nullptr)) return 0;
            } else {
                if (!add_pred(PR_FILTER, -1, (int)nd.aux, 0, 0, nullptr)) return 0;
            }
        } else {
            if (nd.multi) return 0;
            int sc = sh.stage_col(m.col_slot[node]);
            if (sc < 0) return 0;
            if (!add_pred(nd.pred == TAGG_PRED_LUT ? PR_LUT : PR_RANGE, sc, -1, nd.u0, nd.u1,
                          nd.pred == TAGG_PRED_LUT ? es.plan->d_blobs[nd.aux] : nullptr))
                return 0;
            pred_const[sp.n_preds - 1] = {nd.u0, nd.u1};
        }
        node++;
    }
    if (node >= n_nodes) return 0;
    for (uint32_t i = node; i < n_nodes; i++)  // narrowing nodes deeper in the tree are not part of the flat shape
        if (m.nodes[i].op == TAGG_OP_FILTER || m.nodes[i].op == TAGG_OP_POST_FILTER) return 0;

    int n_rgroups = 0, n_bgroups = 0, bucket_mode = BK_NONE, bucket_scope = -1;
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
            sp.root_count_acc[sp.n_root_counts++] = (uint64_t*)(es.arena + es.slots[m.slot_of[mem]].off_acc);
        } else if (nd.op == TAGG_OP_SUM || nd.op == TAGG_OP_MIN || nd.op == TAGG_OP_MAX) {
            if (!add_fold(es, sh, sp.rgroups, n_rgroups, ST_MAXRG, (int)mem)) return 0;
        } else if (nd.op == TAGG_OP_TERMS || nd.op == TAGG_OP_HISTOGRAM) {
            if (bucket_mode != BK_NONE || nd.multi) return 0;
            bucket_scope = m.own_scope[mem];
            const ScopeLayout& L = es.scopes[bucket_scope];
            if (L.mode != SCOPE_DENSE) return 0;
            bucket_mode = nd.op == TAGG_OP_TERMS ? BK_TERMS : BK_HIST;
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
                    sp.bcount_acc[sp.n_bcounts++] = (uint64_t*)(es.arena + es.slots[m.slot_of[lf]].off_acc);
                } else if (ln.op == TAGG_OP_SUM || ln.op == TAGG_OP_MIN || ln.op == TAGG_OP_MAX) {
                    if (!add_fold(es, sh, sp.bgroups, n_bgroups, ST_MAXBG, (int)lf)) return 0;
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
        uint32_t maxnb = 0;
        for (auto& hs : es.hsegs) maxnb = std::max(maxnb, hs.cols[sh.staged[c]].num_bits);
        sp.soff_col[c] = off;
        off += (ST_TILE / 8) * maxnb + 16;
    }
    off = (off + 127) & ~127u;
    sp.soff_bits = off;
    off += 256 * ST_MAXBITS;
    sp.stage_bytes = (off + 127) & ~127u;
    size_t smem_bytes = (size_t)ST_STAGES * sp.stage_bytes + ST_WARPS * ST_DOCS_PER_WARP * 2 + ST_STAGES * 8 + 2 * ST_STAGES * 4 + 64;
    if (smem_bytes > 200 * 1024) return 0;

    // per-segment descriptors
    std::vector<SegDesc> descs(nseg);
    uint64_t tiles_total = 0;
    bool narrowing = sp.n_preds > 0;
    for (size_t i = 0; i < nseg; i++) {
        const DevSegment& hs = es.hsegs[i];
        SegDesc& d = descs[i];
        memset(&d, 0, sizeof(d));
        if (tiles_total > 0xffffffffull) return 0;
        d.tile_begin = (uint32_t)tiles_total;
        tiles_total += ((uint64_t)hs.max_doc + ST_TILE - 1) / ST_TILE;
        d.max_doc = hs.max_doc;
        for (int c = 0; c < sp.n_cols; c++) {
            const DevColumn& col = hs.cols[sh.staged[c]];
            d.col_ptr[c] = (const uint8_t*)col.words;
            d.nb[c] = col.num_bits;
            d.minv[c] = col.min_value;
        }
        if (hs.main.kind == DS_BITSET) { d.flags |= SF_MAIN_BITS; d.bits_ptr[0] = (const uint8_t*)hs.main.words; narrowing = true; }
        if (hs.has_deletes) { d.flags |= SF_DELETES; d.bits_ptr[1] = (const uint8_t*)hs.deleted; narrowing = true; }
        for (int pi = 0; pi < sp.n_preds; pi++) {
            int src = pred_src[pi];
            if (src == -2) { d.pred_lo[pi] = hs.main.lo; d.pred_hi[pi] = hs.main.hi; }
            else if (src == -1) { d.pred_lo[pi] = pred_const[pi].first; d.pred_hi[pi] = pred_const[pi].second; }
            else {
                const DevDocset& fd = hs.filters[src];
                if (sp.pred_type[pi] == PR_RANGE) { d.pred_lo[pi] = fd.lo; d.pred_hi[pi] = fd.hi; }
                else if (fd.kind == DS_BITSET) { d.flags |= SF_PRED_BITS0 << pi; d.bits_ptr[2 + pi] = (const uint8_t*)fd.words; }
                else if (fd.kind != DS_ALL) d.flags |= SF_PRED_NONE0 << pi;
            }
        }
    }
    if (tiles_total > 0xffffffffull) return 0;
    sp.n_segs = (uint32_t)nseg;
    sp.n_tiles = (uint32_t)tiles_total;

    // Option flags of the bucket slots coincide with bucket existence in the flat shape: alias them
    if (bucket_mode != BK_NONE) {
        for (size_t k = 0; k < es.slots.size(); k++)
            if (m.scope_of[m.slot_node[k]] == bucket_scope) es.slots[k].off_seen = es.scopes[bucket_scope].off_present;
    }
    es.path_used = 2;
    if (sp.n_tiles == 0) return 1;
    SegDesc* d_descs = nullptr;
    if (cudaMallocAsync((void**)&d_descs, nseg * sizeof(SegDesc), es.st) != cudaSuccess) return -tagg_fail(TAGG_ERR_OOM, "segment table allocation failed");
    es.temps.push_back(d_descs);
    if (cudaMemcpyAsync(d_descs, descs.data(), nseg * sizeof(SegDesc), cudaMemcpyHostToDevice, es.st) != cudaSuccess)
        return -tagg_fail(TAGG_ERR_CUDA, "segment table upload failed");
    sp.segs = d_descs;
    uint8_t* present = sp.present;
    if (bucket_mode != BK_NONE && sp.n_bcounts > 0) sp.present = nullptr;  // derived from the counts below

    // compaction pays when documents are filtered out; with nothing narrowing the stream it is pure overhead
    int nrg_t = n_rgroups <= 1 ? n_rgroups : ST_MAXRG;
    stream_fn fn = pick_kernel(bucket_mode, n_bgroups, nrg_t, narrowing);
    static std::mutex attr_mu;
    static std::vector<stream_fn> attr_done;
    {
        std::lock_guard<std::mutex> g(attr_mu);
        if (std::find(attr_done.begin(), attr_done.end(), fn) == attr_done.end()) {
            if (cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
                return -tagg_fail(TAGG_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
            attr_done.push_back(fn);
        }
    }
    int per_sm = (int)std::min<size_t>(2048 / ST_THREADS, (227 * 1024) / (smem_bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)es.ctx->sm_count * per_sm, sp.n_tiles);
    fn<<<grid, ST_THREADS, smem_bytes, es.st>>>(sp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "k_stream launch failed: %s", cudaGetErrorString(e));
    es.ctx->launches++;
    es.n_launches++;
    if (bucket_mode != BK_NONE && sp.n_bcounts > 0) {
        uint64_t n = es.scopes[bucket_scope].capacity;
        k_present_from_counts<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 1024), 256, 0, es.st>>>(sp.bcount_acc[0], present, n);
        es.ctx->launches++;
        es.n_launches++;
    }
    return 1;
}

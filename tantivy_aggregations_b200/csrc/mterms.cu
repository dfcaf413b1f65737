// mterms.cu — K5: terms buckets keyed by a MULTI-VALUED field (terms.rs:172-179) or by a key domain too
// wide for a dense table (global open-addressing spill table), with count / sum / min / max leaves on
// single- or multi-valued columns (sum.rs:131-140, minmax.rs:135-145) — BASELINE config C4.
//
// The reference touches a bucket once per VALUE OCCURRENCE of the key field and lets the nested leaves
// collect the DOCUMENT each time, i.e. sum_agg_f64s under terms_agg_u64s adds all of the document's
// values once per key occurrence.  Here a sub-block of 256 threads (up to 4 per CTA, independent, named barriers) owns
// a tile of 1024 documents:
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
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "exec.h"
#include "narrow.cuh"

#define MT_SUB_THREADS 256
#define MT_MAXSUB 4
#define MT_TILE 1024     // documents per sub-block tile
#define MT_TILE_BITS 10
#define MT_CHUNK 4096    // key occurrences expanded at a time
#define MT_STAGE (MT_CHUNK * 2 / 8)  // leaf values staged at a time (the expansion buffer, as 64-bit codes)
#define MT_MAXGROUPS 2
#define MT_MAXPRED NARROW_MAXPRED
#define MT_MAXCOUNTS 2
#define MT_U 4           // key occurrences in flight per thread (8 spills)
#define MT_DU (MT_TILE / MT_SUB_THREADS)  // documents per thread in the doc phase
#define MT_CACHE_LOG 9   // hot-key front: 2^9 entries per CTA (10.5 KB: four sub-blocks and the 1 M-key bitmap still fit)
#define MT_CACHE_EMPTY 0xFFFFFFFFFFFFFFFFull

enum { MO_SUM = 1, MO_MIN = 2, MO_MAX = 4 };
// Option flags ("the leaf saw a value in this bucket", sum.rs:97-101 / minmax.rs:99-105) without a scattered
// global access per occurrence:
//   SEEN_BUCKET   single-valued leaf: Some exactly where the bucket exists (aliased to the bucket-existence flags)
//   SEEN_DERIVED  read off an accumulator after the pass (k_mterms_fixup): a min / max cell that left its identity,
//                 or an f64 sum cell that left -0.0 (every f64 sum cell starts at -0.0: x + -0.0 == x for every x, and
//                 only contributions that ARE the identity need an explicit flag store)
//   SEEN_EXPLICIT check-and-set per occurrence (integer sums only: every bit pattern is a legitimate sum)
enum { SEEN_BUCKET = 0, SEEN_DERIVED = 1, SEEN_EXPLICIT = 2 };
// bucket existence: from the bucket counts after the pass | CTA bitmap in shared memory flushed at the end |
// check-and-set per occurrence (the bitmap does not fit) | the hash table's own slot states
enum { PRESENT_COUNTS = 0, PRESENT_BITMAP = 1, PRESENT_EXPLICIT = 2, PRESENT_HASH = 3 };
#define NEG_ZERO_BITS 0x8000000000000000ull

struct MGroup {
    int32_t col;      // device column slot (multi: idx column, values at col + 1)
    uint32_t kind, multi, ops;
    uint64_t *acc_sum, *acc_min, *acc_max;
    uint8_t* seen;    // Option flags of the group's slots (one array, aliased by all of them)
    uint32_t seen_mode;   // SEEN_*: how the Option flags of this group are produced
    uint32_t derive_op;   // SEEN_DERIVED: the op whose accumulator tells (MO_MIN / MO_MAX: cell != 0; MO_SUM f64: cell != -0.0)
    uint32_t soff_sum, soff_min, soff_max;  // per-document folded contribution, offsets inside a sub-block's shared block
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
    uint32_t present_mode;   // PRESENT_*
    uint32_t bitmap_bytes;   // PRESENT_BITMAP: CTA-private bucket-existence bitmap in shared memory
    uint32_t* front_stat;    // CACHE: [probes, hits] over every sub-block's two measured tiles — the host remembers the verdict
    uint32_t cache_bytes;    // hot-key front (CACHE instantiations): [keys u64][sums u64][counts u32][touched u8] x 2^MT_CACHE_LOG
    uint32_t sub_bytes;      // shared bytes per sub-block
    uint32_t soff_koff, soff_docof, soff_flags, soff_cols;
};

__device__ __forceinline__ void named_bar(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Column descriptors of the current segment, cached in the sub-block's shared memory (slot 0: key offsets,
// 1: key values, 2 + 2g / 3 + 2g: offsets / values of leaf group g).  Both words of a value are always
// fetched (the allocation is padded, dev.cuh), so the two loads are independent and nothing branches on
// the straddle.
struct ColS {
    const uint64_t* words;
    uint64_t minv, mask;
    uint32_t nb, pad;
};
#define MT_NCOLS (2 + 2 * MT_MAXGROUPS)
__device__ __forceinline__ uint64_t cget(const ColS& c, uint64_t i) {
    const uint64_t bit = i * c.nb;
    if (c.nb <= 32) {  // narrow columns (keys, offsets): 32-bit words and one funnel shift
        const uint32_t* w32 = (const uint32_t*)c.words + (bit >> 5);
        const uint32_t lo = __ldg(w32), hi = __ldg(w32 + 1);
        return (uint64_t)(__funnelshift_r(lo, hi, (uint32_t)bit & 31u) & (uint32_t)c.mask) + c.minv;
    }
    const uint64_t w = bit >> 6;
    const uint32_t sh = (uint32_t)bit & 63u;
    const uint64_t lo = __ldg(c.words + w), hi = __ldg(c.words + w + 1);
    const uint64_t v = (lo >> sh) | ((hi << 1) << (63u - sh));
    return (v & c.mask) + c.minv;
}
// ask L2 for the packed bytes of values [lo, hi) of a column (the sub-block's next tile)
__device__ __forceinline__ void l2_prefetch_values(const ColS& c, uint64_t lo, uint64_t hi) {
    if (hi <= lo || c.nb == 0) return;
    const uint64_t b0 = (lo * c.nb) >> 3, b1 = (hi * c.nb + 7) >> 3;
    const uint8_t* p0 = (const uint8_t*)c.words + (b0 & ~(uint64_t)15);
    uint64_t bytes = (b1 - (b0 & ~(uint64_t)15) + 15) & ~(uint64_t)15;
    if (bytes > (1u << 20)) bytes = 1u << 20;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p0), "r"((uint32_t)bytes) : "memory");
}
// Entries first, first + 1, ... of a column relative to `first`: 32-bit arithmetic per entry (i * nb stays small inside a tile).
// Rel32: packed DELTAS of a narrow column (nb <= 32) — offsets columns, whose min_value cancels in differences.
struct Rel32 { const uint32_t* wp; uint32_t sh0, nb, mask; };
__device__ __forceinline__ Rel32 rel32_at(const ColS& c, uint64_t first) {
    const uint64_t bit0 = first * c.nb;
    Rel32 r;
    r.wp = (const uint32_t*)c.words + (bit0 >> 5); r.sh0 = (uint32_t)bit0 & 31u; r.nb = c.nb; r.mask = (uint32_t)c.mask;
    return r;
}
__device__ __forceinline__ uint32_t rel32_delta(const Rel32& r, uint32_t i) {
    const uint32_t bb = r.sh0 + i * r.nb;
    const uint32_t* w = r.wp + (bb >> 5);
    return __funnelshift_r(__ldg(w), __ldg(w + 1), bb & 31u) & r.mask;
}
// RelCol: codes of a column of any width
struct RelCol { const uint64_t* wp; uint64_t mask, minv; uint32_t sh0, nb; };
__device__ __forceinline__ RelCol rel_at(const ColS& c, uint64_t first) {
    const uint64_t bit0 = first * c.nb;
    RelCol r;
    r.wp = c.words + (bit0 >> 6); r.sh0 = (uint32_t)bit0 & 63u; r.nb = c.nb; r.mask = c.mask; r.minv = c.minv;
    return r;
}
__device__ __forceinline__ uint64_t rel_get(const RelCol& r, uint32_t j) {
    const uint32_t bb = r.sh0 + j * r.nb;
    const uint64_t* w = r.wp + (bb >> 6);
    const uint32_t sh = bb & 63u;
    const uint64_t lo = __ldg(w), hi = __ldg(w + 1);
    return (((lo >> sh) | ((hi << 1) << (63u - sh))) & r.mask) + r.minv;
}
__device__ __forceinline__ void cols_load(ColS* dst, const DevColumn& c) {
    dst->words = c.words; dst->minv = c.min_value; dst->mask = c.mask; dst->nb = c.num_bits; dst->pad = 0;
}

// CACHE: the shared-memory hash front of north_star (4) — a CTA-private table of (key -> partial count / partial sum), first
// come first served, in front of the global table: a hot key of a skewed distribution is folded on chip and reaches the
// global table (where same-address atomics serialise in L2: 3 % of 10^9 updates on ONE cell took 48 ms) once per CTA.
// Keys that find their slot taken go straight to the global table.  For shapes with at most one count and one sum-only
// leaf group.
// OPS0 >= 0: the op mask of leaf group 0, compiled in (the sum-only shape of C4 drops the min / max accumulators of the doc
// phase: the kernel sits at the 64-register cap and every spill shows — 340 -> 72 bytes of spill stores, 9.6 -> 8.2 ms)
template <bool DENSE, int NG, int NC, bool CACHE, int OPS0 = -1>
__global__ void __launch_bounds__(MT_SUB_THREADS * MT_MAXSUB, 1) k_mterms(const __grid_constant__ MParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t tid = threadIdx.x, sub = tid / MT_SUB_THREADS, st = tid % MT_SUB_THREADS;
    const uint32_t n_sub = blockDim.x / MT_SUB_THREADS;
    uint32_t* bitmap = (uint32_t*)smem;
    uint64_t* ckey = (uint64_t*)(smem + p.bitmap_bytes);
    uint64_t* csum = ckey + (1u << MT_CACHE_LOG);
    uint32_t* ccnt = (uint32_t*)(csum + (1u << MT_CACHE_LOG));
    uint8_t* ctouch = (uint8_t*)(ccnt + (1u << MT_CACHE_LOG));
    uint8_t* base = smem + p.bitmap_bytes + p.cache_bytes + (size_t)sub * p.sub_bytes;
    uint32_t* koff = (uint32_t*)(base + p.soff_koff);    // MT_TILE + 1 key offsets relative to the tile's first key
    uint16_t* docof = (uint16_t*)(base + p.soff_docof);  // MT_CHUNK tile-local document indices
    uint8_t* flags = base + p.soff_flags;                // per document: bit 0 matched, bit 1 + g contributes to group g
    ColS* cs = (ColS*)(base + p.soff_cols);
    const bool has_bitmap = DENSE && p.present_mode == PRESENT_BITMAP;
    if (has_bitmap)
        for (uint32_t i = tid; i < p.bitmap_bytes / 4; i += blockDim.x) bitmap[i] = 0;
    if (CACHE)
        for (uint32_t i = tid; i < (1u << MT_CACHE_LOG); i += blockDim.x) {
            ckey[i] = MT_CACHE_EMPTY;
            csum[i] = (NG > 0 && p.groups[0].kind == TAGG_F64) ? NEG_ZERO_BITS : 0ull;
            ccnt[i] = 0;
            ctouch[i] = 0;
        }
    __syncthreads();
    uint32_t cur_seg = 0xffffffffu, seg = 0;
    // the front pays only for skewed keys: every sub-block measures its hit rate over its first two tiles and stops
    // probing below 1/64 (uniform keys over a large domain: ~10 % of the kernel for nothing); cached entries stay valid
    bool use_cache = CACHE;
    uint32_t n_probe = 0, n_hit = 0, tiles_done = 0;
    uint32_t* cstat = (uint32_t*)(ctouch + (1u << MT_CACHE_LOG)) + 2 * sub;  // [probes, hits] of this sub-block
    if (CACHE && st == 0) { cstat[0] = 0; cstat[1] = 0; }
    const uint64_t dom_min = p.scope.dom_min, dom_size = p.scope.dom_size;

    for (uint64_t tile = (uint64_t)blockIdx.x * n_sub + sub; tile < p.n_tiles; tile += (uint64_t)gridDim.x * n_sub) {
        while (seg + 1 < p.n_segs && __ldg(p.tile_begin + seg + 1) <= (uint32_t)tile) seg++;
        const DevSegment& S = p.segs[seg];
        if (seg != cur_seg) {  // uniform over the sub-block: refresh the cached column descriptors
            named_bar(1 + sub, MT_SUB_THREADS);
            if (st == 0) cols_load(cs + 0, S.cols[p.key_col]);
            if (st == 1) cols_load(cs + 1, S.cols[p.key_multi ? p.key_col + 1 : p.key_col]);
            if (st >= 2 && st < 2 + 2 * NG) {
                const MGroup& G = p.groups[(st - 2) >> 1];
                cols_load(cs + st, S.cols[G.multi ? G.col + ((st - 2) & 1) : G.col]);
            }
            named_bar(1 + sub, MT_SUB_THREADS);
            cur_seg = seg;
        }
        const uint32_t d0 = ((uint32_t)tile - __ldg(p.tile_begin + seg)) * MT_TILE;
        const uint32_t nd = min((uint32_t)MT_TILE, S.max_doc - d0);
        const bool koff32 = p.key_multi && cs[0].nb <= 32;  // narrow offsets: every column of < 2^32 values
        const uint32_t kd0 = koff32 ? rel32_delta(rel32_at(cs[0], d0), 0) : 0u;
        const uint64_t kbase = koff32 ? (uint64_t)kd0 + cs[0].minv : p.key_multi ? cget(cs[0], d0) : (uint64_t)d0;
        const bool plain = S.main.kind == DS_ALL && !S.has_deletes && p.n_preds == 0;
        // ---- L2 prefetch of the sub-block's NEXT tile: fixed-position slices (offset columns) now, the
        //      value slices (whose position depends on the offsets at the tile's ends) after the doc phase
        const uint64_t ntile = tile + (uint64_t)gridDim.x * n_sub;
        uint64_t pf_lo = 0, pf_hi = 0;  // lanes 1 / 3 + 2g of warp 0: value range of the next tile
        int pf_col = -1;
        if (st < 2 + 2 * NG && ntile < p.n_tiles && ntile < __ldg(p.tile_begin + seg + 1)) {
            const uint32_t nd0 = ((uint32_t)ntile - __ldg(p.tile_begin + seg)) * MT_TILE;
            const uint32_t nnd = min((uint32_t)MT_TILE, S.max_doc - nd0);
            const bool is_vals = st & 1;
            const bool multi = st < 2 ? p.key_multi != 0 : p.groups[(st - 2) >> 1].multi != 0;
            if (!is_vals) {
                if (multi) l2_prefetch_values(cs[st], nd0, (uint64_t)nd0 + nnd + 1);
            } else {
                pf_col = (int)st;
                if (multi) { pf_lo = cget(cs[st - 1], nd0); pf_hi = cget(cs[st - 1], (uint64_t)nd0 + nnd); }
                else { pf_lo = nd0; pf_hi = (uint64_t)nd0 + nnd; }
            }
        }
        // ---- doc phase: MT_DU documents per thread in flight --------------------------------------------
        // Offsets (key idx, leaf idx) are unpacked by consecutive threads from consecutive packed entries — coalesced, in
        // 32-bit arithmetic relative to the tile's first entry (the column's min_value cancels in the differences).  A
        // multi-valued leaf is then folded from shared memory: the tile's value slice [idx[d0], idx[d0 + nd]) is unpacked
        // one value per thread (coalesced again) into the chunk buffer, and each document folds its own range there.  The
        // thread-per-document walk idx -> vals it replaces was a chain of dependent, scattered L2 loads: 40 % of the
        // kernel's stall samples (profiles/r2_ncu_hot_C4.txt).
        {
            uint32_t fl[MT_DU];
            const Rel32 kr = rel32_at(cs[0], d0);
#pragma unroll
            for (int u = 0; u < MT_DU; u++) {
                const uint32_t i = st + u * MT_SUB_THREADS;
                fl[u] = 0;
                if (i >= nd) continue;
                koff[i] = !p.key_multi ? i : koff32 ? rel32_delta(kr, i) - kd0 : (uint32_t)(cget(cs[0], d0 + i) - kbase);
                fl[u] = (plain || doc_matches(S, p.preds, p.n_preds, d0 + i)) ? 1u : 0u;
            }
            if (st == 0) koff[nd] = !p.key_multi ? nd : koff32 ? rel32_delta(kr, nd) - kd0 : (uint32_t)(cget(cs[0], (uint64_t)d0 + nd) - kbase);
#pragma unroll
            for (int g = 0; g < NG; g++) {
                const MGroup& G = p.groups[g];
                const uint32_t gops = (OPS0 >= 0 && g == 0) ? (uint32_t)OPS0 : G.ops;
                const ColS& oc = cs[2 + 2 * g];
                const ColS& vc = cs[3 + 2 * g];
                uint64_t sum[MT_DU], mn[MT_DU], mx[MT_DU];  // min in max-form (~code), like the arena
                bool has[MT_DU];
#pragma unroll
                for (int u = 0; u < MT_DU; u++) { sum[u] = G.kind == TAGG_F64 ? NEG_ZERO_BITS : 0ull; mn[u] = 0; mx[u] = 0; has[u] = false; }  // f64 sums fold from -0.0
                auto fold = [&](int u, uint64_t code) {
                    if (gops & MO_SUM) {
                        if (G.kind == TAGG_F64) sum[u] = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)sum[u]), code_to_f64(code)));
                        else sum[u] += code_to_bits(G.kind, code);
                    }
                    if (gops & MO_MIN) mn[u] = max(mn[u], ~code);
                    if (gops & MO_MAX) mx[u] = max(mx[u], code);
                };
                if (G.multi && oc.nb <= 32) {
                    // staged: the ranges live in the group's first contribution array until they are in registers
                    uint32_t* voff = (uint32_t*)(base + ((gops & MO_SUM) ? G.soff_sum : (gops & MO_MIN) ? G.soff_min : G.soff_max));
                    uint64_t* stage = (uint64_t*)docof;  // MT_STAGE codes; the expansion buffer is idle during the doc phase
                    const Rel32 vr = rel32_at(oc, d0);
                    const uint32_t vd0 = rel32_delta(vr, 0);
                    const uint64_t vbase = (uint64_t)vd0 + oc.minv;  // index of the tile's first value
#pragma unroll
                    for (int u = 0; u < MT_DU; u++) {
                        const uint32_t i = st + u * MT_SUB_THREADS;
                        if (i < nd) voff[i] = rel32_delta(vr, i) - vd0;
                    }
                    if (st == 0) voff[nd] = rel32_delta(vr, nd) - vd0;
                    named_bar(1 + sub, MT_SUB_THREADS);
                    uint32_t ga[MT_DU], ge[MT_DU];
#pragma unroll
                    for (int u = 0; u < MT_DU; u++) {
                        const uint32_t i = st + u * MT_SUB_THREADS;
                        ga[u] = ge[u] = 0;
                        if (i < nd && fl[u]) { ga[u] = voff[i]; ge[u] = voff[i + 1]; }
                        has[u] = ge[u] > ga[u];
                    }
                    const uint32_t nv = voff[nd];
                    named_bar(1 + sub, MT_SUB_THREADS);  // every range is in registers: voff's array may be written, the buffer filled
                    for (uint32_t vb = 0; vb < nv; vb += MT_STAGE) {
                        const uint32_t cn = min((uint32_t)MT_STAGE, nv - vb);
                        const RelCol rv = rel_at(vc, vbase + vb);
                        for (uint32_t j = st; j < cn; j += MT_SUB_THREADS) stage[j] = rel_get(rv, j);
                        named_bar(1 + sub, MT_SUB_THREADS);
                        uint32_t lo[MT_DU], hi[MT_DU];
#pragma unroll
                        for (int u = 0; u < MT_DU; u++) { lo[u] = max(ga[u], vb); hi[u] = min(ge[u], vb + cn); }
                        for (uint32_t r = 0;; r++) {  // round r: the r-th value (inside this chunk) of each of the thread's documents
                            uint64_t code[MT_DU];
                            bool live[MT_DU], any = false;
#pragma unroll
                            for (int u = 0; u < MT_DU; u++) {
                                live[u] = lo[u] + r < hi[u];
                                code[u] = live[u] ? stage[lo[u] + r - vb] : 0;
                                any = any || live[u];
                            }
                            if (!any) break;
#pragma unroll
                            for (int u = 0; u < MT_DU; u++)
                                if (live[u]) fold(u, code[u]);
                        }
                        named_bar(1 + sub, MT_SUB_THREADS);
                    }
                } else {
                    uint64_t ga[MT_DU], ge[MT_DU];
#pragma unroll
                    for (int u = 0; u < MT_DU; u++) {
                        const uint32_t doc = d0 + st + u * MT_SUB_THREADS;
                        ga[u] = ge[u] = 0;
                        if (!fl[u]) continue;
                        ga[u] = doc; ge[u] = (uint64_t)doc + 1;
                        if (G.multi) { ga[u] = cget(oc, doc); ge[u] = cget(oc, (uint64_t)doc + 1); }
                        has[u] = ge[u] > ga[u];
                    }
                    for (uint64_t r = 0;; r++) {  // round r: the r-th value of each of the thread's documents
                        uint64_t code[MT_DU];
                        bool live[MT_DU], any = false;
#pragma unroll
                        for (int u = 0; u < MT_DU; u++) {
                            live[u] = ga[u] + r < ge[u];
                            code[u] = live[u] ? cget(vc, ga[u] + r) : 0;
                            any = any || live[u];
                        }
                        if (!any) break;
#pragma unroll
                        for (int u = 0; u < MT_DU; u++)
                            if (live[u]) fold(u, code[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < MT_DU; u++) {
                    if (has[u]) {
                        const uint32_t i = st + u * MT_SUB_THREADS;
                        fl[u] |= 2u << g;
                        if (gops & MO_SUM) ((uint64_t*)(base + G.soff_sum))[i] = sum[u];
                        if (gops & MO_MIN) ((uint64_t*)(base + G.soff_min))[i] = mn[u];
                        if (gops & MO_MAX) ((uint64_t*)(base + G.soff_max))[i] = mx[u];
                        // a contribution equal to the accumulator's identity leaves no trace in the cell: the value phase sets
                        // the Option flag explicitly for these documents (SEEN_DERIVED)
                        const uint64_t own = (OPS0 == MO_SUM || G.derive_op == MO_SUM) ? sum[u] : G.derive_op == MO_MIN ? mn[u] : mx[u];
                        if (G.seen_mode == SEEN_DERIVED && own == (G.derive_op == MO_SUM ? NEG_ZERO_BITS : 0ull)) fl[u] |= 8u << g;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < MT_DU; u++) {
                const uint32_t i = st + u * MT_SUB_THREADS;
                if (i < nd) flags[i] = (uint8_t)fl[u];
            }
        }
        if (pf_col >= 0) l2_prefetch_values(cs[pf_col], pf_lo, pf_hi);
        named_bar(1 + sub, MT_SUB_THREADS);
        const uint32_t nk = koff[nd];
        for (uint32_t cbase = 0; cbase < nk; cbase += MT_CHUNK) {
            const uint32_t cn = min((uint32_t)MT_CHUNK, nk - cbase);
            // ---- expand: tile-local document index of every key occurrence of the chunk ------------------
            // (the document's flag bits ride in the top bits of the entry: one shared load per occurrence)
            for (uint32_t i = st; i < nd; i += MT_SUB_THREADS) {
                uint32_t lo = max(koff[i], cbase), hi = min(koff[i + 1], cbase + cn);
                const uint16_t e = (uint16_t)(i | ((uint32_t)flags[i] << MT_TILE_BITS));
                for (uint32_t v = lo; v < hi; v++) docof[v - cbase] = e;
            }
            named_bar(1 + sub, MT_SUB_THREADS);
            const uint32_t knb = cs[1].nb, kmask = (uint32_t)cs[1].mask;
            const bool k32 = knb >= 1 && knb <= 32;
            const uint64_t kminv = cs[1].minv, kbit0 = (kbase + cbase) * knb;
            const uint32_t* kwp = (const uint32_t*)cs[1].words + (kbit0 >> 5);
            const uint32_t ksh0 = (uint32_t)kbit0 & 31u;
            // ---- value phase: one thread per key occurrence (terms.rs:172-179), MT_U occurrences in flight ---
            for (uint32_t v0 = st; v0 < cn; v0 += MT_SUB_THREADS * MT_U) {
                uint32_t di[MT_U], f[MT_U], b[MT_U];
                uint64_t key[MT_U];
#pragma unroll
                for (int u = 0; u < MT_U; u++) {
                    const uint32_t v = min(v0 + u * MT_SUB_THREADS, cn - 1);
                    const uint32_t e = docof[v];
                    di[u] = e & (MT_TILE - 1);
                    f[u] = v0 + u * MT_SUB_THREADS < cn ? e >> MT_TILE_BITS : 0u;
                    if (k32) {  // narrow keys: 32-bit arithmetic relative to the chunk's first key
                        const uint32_t bb = ksh0 + v * knb;
                        const uint32_t* w = kwp + (bb >> 5);
                        key[u] = (uint64_t)(__funnelshift_r(__ldg(w), __ldg(w + 1), bb & 31u) & kmask) + kminv;
                    } else {
                        key[u] = cget(cs[1], kbase + cbase + v);
                    }
                }
                if (CACHE && use_cache) {
                    // hot-key front: a key that owns (or can claim) its slot is folded in shared memory
#pragma unroll
                    for (int u = 0; u < MT_U; u++) {
                        if (!f[u]) continue;
                        if (DENSE && key[u] - dom_min >= dom_size) { f[u] = 0; continue; }
                        n_probe++;
                        // two slots per key (first come, first served in each): a hot key goes without a slot only if colder
                        // keys took BOTH before its first occurrence in this CTA — every such CTA sends all of the key's
                        // updates to one global cell, where they serialise
                        uint32_t slot = ((uint32_t)key[u] * 0x9E3779B1u) >> (32 - MT_CACHE_LOG);
                        uint64_t k = ckey[slot];
                        bool claimed = false;  // (a claim is not reuse: it does not count as a hit)
                        if (k == MT_CACHE_EMPTY) {
                            k = atomicCAS((unsigned long long*)(ckey + slot), (unsigned long long)MT_CACHE_EMPTY, (unsigned long long)key[u]);
                            if (k == MT_CACHE_EMPTY) { k = key[u]; claimed = true; }
                        }
                        if (k != key[u]) {
                            slot = ((uint32_t)(key[u] >> 7) * 0x85EBCA6Bu + (uint32_t)key[u] * 0xC2B2AE35u) >> (32 - MT_CACHE_LOG);
                            k = ckey[slot];
                            if (k == MT_CACHE_EMPTY) {
                                k = atomicCAS((unsigned long long*)(ckey + slot), (unsigned long long)MT_CACHE_EMPTY, (unsigned long long)key[u]);
                                if (k == MT_CACHE_EMPTY) { k = key[u]; claimed = true; }
                            }
                            if (k != key[u]) continue;  // both slots belong to other keys: global table
                        }
                        n_hit += claimed ? 0u : 1u;
                        if (NC > 0) atomicAdd(ccnt + slot, 1u);
                        if (NG > 0 && ((f[u] >> 1) & 1u)) {
                            const uint64_t v = ((const uint64_t*)(base + p.groups[0].soff_sum))[di[u]];
                            if (p.groups[0].kind == TAGG_F64) atomicAdd((double*)(csum + slot), __longlong_as_double((long long)v));
                            else atomicAdd((unsigned long long*)(csum + slot), (unsigned long long)v);
                            if (!ctouch[slot]) ctouch[slot] = 1;
                        }
                        f[u] = 0;  // done
                    }
                }
                if (DENSE) {
#pragma unroll
                    for (int u = 0; u < MT_U; u++) {
                        const uint64_t rel = key[u] - dom_min;  // below the domain: wraps to a huge value
                        if (rel >= dom_size) f[u] = 0;
                        b[u] = f[u] ? (uint32_t)rel : 0u;
                        if (p.present_mode == PRESENT_BITMAP) {
                            if (f[u] && !((bitmap[b[u] >> 5] >> (b[u] & 31)) & 1u)) atomicOr(bitmap + (b[u] >> 5), 1u << (b[u] & 31));
                        } else if (p.present_mode == PRESENT_EXPLICIT) {
                            if (f[u] && !p.scope.present[b[u]]) p.scope.present[b[u]] = 1;
                        }
                    }
                } else {
                    // the home slots of all MT_U keys are read at once (independent L2 loads); a key found there is done — at
                    // load <= 3/4 most are — and only the rest walk the probe sequence one dependent load at a time
                    const uint64_t hmask = p.scope.capacity - 1;
                    uint64_t home[MT_U], hkey[MT_U];
#pragma unroll
                    for (int u = 0; u < MT_U; u++) {
                        home[u] = hash_key(key[u], 0) >> __clzll(hmask);
                        hkey[u] = f[u] ? __ldcg(p.scope.keys + home[u]) : 0ull;
                    }
#pragma unroll
                    for (int u = 0; u < MT_U; u++) {
                        b[u] = 0;
                        if (!f[u]) continue;
                        // (an all-zero cell is ambiguous — empty or key 0 — and takes the state-checked path)
                        b[u] = (hkey[u] != 0 && hkey[u] == key[u]) ? (uint32_t)home[u] : scope_lookup(p.overflow, p.scope, 0, key[u]);
                        if (b[u] == INVALID_BUCKET) { f[u] = 0; b[u] = 0; }
                    }
                }
                // Option flags that cannot be read off an accumulator afterwards
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const MGroup& G = p.groups[g];
                const uint32_t gops = (OPS0 >= 0 && g == 0) ? (uint32_t)OPS0 : G.ops;
                    if (G.seen_mode == SEEN_EXPLICIT) {
#pragma unroll
                        for (int u = 0; u < MT_U; u++)
                            if (((f[u] >> (1 + g)) & 1u) && !G.seen[b[u]]) G.seen[b[u]] = 1;
                    } else if (G.seen_mode == SEEN_DERIVED) {  // a contribution equal to the identity leaves no trace in the
                        // cell (min / max: the smallest / largest code; f64 sum: a document whose values are all -0.0): the
                        // doc phase marked those documents (flag bit 3 + g)
#pragma unroll
                        for (int u = 0; u < MT_U; u++)
                            if ((f[u] >> (3 + g)) & 1u) G.seen[b[u]] = 1;
                    }
                }
#pragma unroll
                for (int u = 0; u < MT_U; u++) {
                    if (NC > 0 && f[u]) atomicAdd((unsigned long long*)(p.count_acc[0] + b[u]), 1ull);
                    if (NC > 1 && f[u]) atomicAdd((unsigned long long*)(p.count_acc[1] + b[u]), 1ull);
                }
#pragma unroll
                for (int g = 0; g < NG; g++) {
                    const MGroup& G = p.groups[g];
                const uint32_t gops = (OPS0 >= 0 && g == 0) ? (uint32_t)OPS0 : G.ops;
                    if (gops & MO_SUM) {
                        const uint64_t* ss = (const uint64_t*)(base + G.soff_sum);
                        if (G.kind == TAGG_F64) {
#pragma unroll
                            for (int u = 0; u < MT_U; u++)
                                if ((f[u] >> (1 + g)) & 1u) atomicAdd((double*)(G.acc_sum + b[u]), __longlong_as_double((long long)ss[di[u]]));
                        } else {
#pragma unroll
                            for (int u = 0; u < MT_U; u++)
                                if ((f[u] >> (1 + g)) & 1u) atomicAdd((unsigned long long*)(G.acc_sum + b[u]), (unsigned long long)ss[di[u]]);
                        }
                    }
                    if (gops & MO_MIN) {
                        const uint64_t* ss = (const uint64_t*)(base + G.soff_min);
#pragma unroll
                        for (int u = 0; u < MT_U; u++) {
                            if (!((f[u] >> (1 + g)) & 1u)) continue;
                            const uint64_t m = ss[di[u]];
                            if (G.acc_min[b[u]] < m && __ldcg(G.acc_min + b[u]) < m) atomicMax((unsigned long long*)(G.acc_min + b[u]), (unsigned long long)m);
                        }
                    }
                    if (gops & MO_MAX) {
                        const uint64_t* ss = (const uint64_t*)(base + G.soff_max);
#pragma unroll
                        for (int u = 0; u < MT_U; u++) {
                            if (!((f[u] >> (1 + g)) & 1u)) continue;
                            const uint64_t m = ss[di[u]];
                            if (G.acc_max[b[u]] < m && __ldcg(G.acc_max + b[u]) < m) atomicMax((unsigned long long*)(G.acc_max + b[u]), (unsigned long long)m);
                        }
                    }
                }
            }
            named_bar(1 + sub, MT_SUB_THREADS);
        }
        if (nk == 0) named_bar(1 + sub, MT_SUB_THREADS);  // (a tile without keys: nobody may still read koff[nd] when the next tile writes it)
        if (CACHE && use_cache && ++tiles_done <= 2) {
            n_probe = __reduce_add_sync(0xffffffffu, n_probe);
            n_hit = __reduce_add_sync(0xffffffffu, n_hit);
            if ((st & 31) == 0) { atomicAdd(cstat, n_probe); atomicAdd(cstat + 1, n_hit); }
            n_probe = n_hit = 0;
            if (tiles_done == 2) {
                named_bar(1 + sub, MT_SUB_THREADS);
                use_cache = cstat[1] * 64u >= cstat[0];
                if (st == 0 && p.front_stat) { atomicAdd(p.front_stat, cstat[0]); atomicAdd(p.front_stat + 1, cstat[1]); }
            }
        }
    }
    if (CACHE) {  // flush the hot-key front: one global update per cached key and CTA
        __syncthreads();
        for (uint32_t i = tid; i < (1u << MT_CACHE_LOG); i += blockDim.x) {
            const uint64_t key = ckey[i];
            if (key == MT_CACHE_EMPTY) continue;
            uint32_t b;
            if (DENSE) {
                b = (uint32_t)(key - dom_min);
                if (p.present_mode == PRESENT_BITMAP) atomicOr(bitmap + (b >> 5), 1u << (b & 31));
                else if (p.present_mode == PRESENT_EXPLICIT) p.scope.present[b] = 1;
            } else {
                b = scope_lookup(p.overflow, p.scope, 0, key);
                if (b == INVALID_BUCKET) continue;
            }
            if (NC > 0 && ccnt[i]) atomicAdd((unsigned long long*)(p.count_acc[0] + b), (unsigned long long)ccnt[i]);
            if (NG > 0 && ctouch[i]) {
                const MGroup& G = p.groups[0];
                const uint64_t v = csum[i];
                if (G.kind == TAGG_F64) atomicAdd((double*)(G.acc_sum + b), __longlong_as_double((long long)v));
                else atomicAdd((unsigned long long*)(G.acc_sum + b), (unsigned long long)v);
                if (G.seen_mode == SEEN_EXPLICIT || (G.seen_mode == SEEN_DERIVED && v == NEG_ZERO_BITS)) G.seen[b] = 1;
            }
        }
    }
    if (has_bitmap) {  // bucket existence, written once per CTA with coalesced stores
        __syncthreads();
        for (uint32_t i = tid; i < (uint32_t)dom_size; i += blockDim.x)
            if ((bitmap[i >> 5] >> (i & 31)) & 1u) p.scope.present[i] = 1;
    }
}

// after the pass: Option flags from the accumulators (every f64 sum cell of the arena starts at -0.0, exec.cu
// layout_arena), bucket existence from the counts
struct MFix {
    uint64_t n;
    uint8_t* present;
    const uint64_t* counts;  // non-null: present[b] = counts[b] != 0
    int32_t n_groups;
    struct { uint64_t* cell; uint8_t* seen; uint32_t derive_op; uint32_t pad; } g[MT_MAXGROUPS];
};
__global__ void k_mterms_fixup(const MFix x) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < x.n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (x.counts) x.present[i] = x.counts[i] != 0;
        for (int g = 0; g < x.n_groups; g++) {
            const uint64_t v = x.g[g].cell[i];
            if (x.g[g].derive_op == MO_SUM) {
                if (v != NEG_ZERO_BITS) x.g[g].seen[i] = 1;
            } else if (v != 0) {
                x.g[g].seen[i] = 1;
            }
        }
    }
}

typedef void (*mterms_fn)(const MParams);
template <bool DENSE, int NG, bool CACHE, int OPS0>
static mterms_fn pick_nc(int nc) {
    switch (nc) {
        case 0: return (mterms_fn)k_mterms<DENSE, NG, 0, CACHE, OPS0>;
        case 1: return (mterms_fn)k_mterms<DENSE, NG, 1, CACHE, OPS0>;
        default: return (mterms_fn)k_mterms<DENSE, NG, 2, false, -1>;
    }
}
template <bool DENSE>
static mterms_fn pick_ng(int ng, int nc, bool cache, bool sum_only) {
    switch (ng) {
        case 0: return cache ? pick_nc<DENSE, 0, true, -1>(nc) : pick_nc<DENSE, 0, false, -1>(nc);
        case 1:
            if (sum_only && nc < 2) return cache ? pick_nc<DENSE, 1, true, MO_SUM>(nc) : pick_nc<DENSE, 1, false, MO_SUM>(nc);
            return cache ? pick_nc<DENSE, 1, true, -1>(nc) : pick_nc<DENSE, 1, false, -1>(nc);
        default: return pick_nc<DENSE, 2, false, -1>(nc);
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
    if (!narrow_chain(es, base.preds, &base.n_preds, &node)) return 0;
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
        for (int g = 0; g < p.n_groups; g++) {
            MGroup& G = p.groups[g];
            if (group_seen[g] == (size_t)-1) {
                G.seen = es.arena + L.off_present;
                G.seen_mode = SEEN_BUCKET;
            } else {
                G.seen = es.arena + group_seen[g];
                G.derive_op = (G.ops & MO_MIN) ? MO_MIN : (G.ops & MO_MAX) ? MO_MAX : (G.kind == TAGG_F64 ? MO_SUM : 0);
                G.seen_mode = G.derive_op ? SEEN_DERIVED : SEEN_EXPLICIT;
            }
        }
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

        // shared-memory plan
        uint32_t off = 0;
        p.soff_koff = off; off += (MT_TILE + 1) * 4; off = (off + 15) & ~15u;
        p.soff_docof = off; off += MT_CHUNK * 2;
        p.soff_flags = off; off += MT_TILE; off = (off + 15) & ~15u;
        p.soff_cols = off; off += MT_NCOLS * (uint32_t)sizeof(ColS);
        for (int g = 0; g < p.n_groups; g++) {
            MGroup& G = p.groups[g];
            if (G.ops & MO_SUM) { G.soff_sum = off; off += MT_TILE * 8; }
            if (G.ops & MO_MIN) { G.soff_min = off; off += MT_TILE * 8; }
            if (G.ops & MO_MAX) { G.soff_max = off; off += MT_TILE * 8; }
        }
        p.sub_bytes = off;
        const size_t SMEM_MAX = 225 * 1024;
        p.bitmap_bytes = 0;
        p.present_mode = dense ? (p.n_counts > 0 ? PRESENT_COUNTS : PRESENT_EXPLICIT) : PRESENT_HASH;
        if (p.present_mode == PRESENT_EXPLICIT) {
            size_t bb = (((size_t)L.dom_size + 31) / 32) * 4;
            bb = (bb + 15) & ~(size_t)15;
            if (bb + 2 * (size_t)p.sub_bytes <= SMEM_MAX) { p.bitmap_bytes = (uint32_t)bb; p.present_mode = PRESENT_BITMAP; }
        }
        // the hot-key front: shapes with at most one count and one sum-only leaf group, when it costs no sub-block
        bool cache = p.n_counts <= 1 && p.n_groups <= 1 && (p.n_groups == 0 || p.groups[0].ops == MO_SUM);
        p.cache_bytes = 0;
        if (cache) {
            const uint32_t cb = ((uint32_t)((8 + 8 + 4 + 1) << MT_CACHE_LOG) + 8 * MT_MAXSUB + 15) & ~15u;
            const size_t without = std::min<size_t>(MT_MAXSUB, (SMEM_MAX - p.bitmap_bytes) / p.sub_bytes);
            const size_t with = p.bitmap_bytes + cb < SMEM_MAX ? std::min<size_t>(MT_MAXSUB, (SMEM_MAX - p.bitmap_bytes - cb) / p.sub_bytes) : 0;
            static const bool no_cache = getenv("TAGG_MT_NOCACHE") != nullptr;  // experiment switch
            if (with >= 1 && with == without && !no_cache) p.cache_bytes = cb; else cache = false;
        }
        // The front's instantiation is the larger kernel (more spills at the 64-register cap): when the previous queries
        // of this plan measured no reuse, the plain instantiation runs (C4 uniform keys: 11.5 -> 9.6 ms); every 32nd query
        // measures again, so a plan whose data turns skewed finds its way back
        static_assert(TAGG_MAX_NODES <= 64, "tagg_plan::mt_front_hint is sized for 64 nodes");
        if (cache) {
            std::lock_guard<std::mutex> g(es.plan->mu);
            uint8_t& hint = es.plan->mt_front_hint[mem];
            if (hint >= 2 && ++hint >= 2 + 32) hint = 0;
            if (hint >= 2) { cache = false; p.cache_bytes = 0; }
        }
        if (cache) { p.front_stat = (uint32_t*)(es.arena + es.off_overflow + 8); es.mt_front_node = (int)mem; }
        uint32_t n_sub = (uint32_t)std::min<size_t>(MT_MAXSUB, (SMEM_MAX - p.bitmap_bytes - p.cache_bytes) / p.sub_bytes);
        if (n_sub < 1) continue;
        const size_t smem_bytes = p.bitmap_bytes + p.cache_bytes + (size_t)n_sub * p.sub_bytes;
        const bool sum_only = p.n_groups == 1 && p.groups[0].ops == MO_SUM;
        mterms_fn fn = dense ? pick_ng<true>(p.n_groups, p.n_counts, cache, sum_only) : pick_ng<false>(p.n_groups, p.n_counts, cache, sum_only);
        {
            static std::mutex attr_mu;
            static std::vector<mterms_fn> attr_done;
            std::lock_guard<std::mutex> g(attr_mu);
            if (std::find(attr_done.begin(), attr_done.end(), fn) == attr_done.end()) {
                if (cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX) != cudaSuccess)
                    return -tagg_fail(TAGG_ERR_CUDA, "cudaFuncSetAttribute(k_mterms) failed: %s", cudaGetErrorString(cudaGetLastError()));
                attr_done.push_back(fn);
            }
        }
        const int threads = (int)n_sub * MT_SUB_THREADS;
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, threads, smem_bytes) != cudaSuccess || per_sm < 1)
            return -tagg_fail(TAGG_ERR_CUDA, "k_mterms does not fit an SM (%zu bytes of shared memory)", smem_bytes);

        if (!es.uploads.empty())
            for (uint32_t c = 0; c < es.n_chunks; c++)
                if (cudaStreamWaitEvent(es.st, es.call->chunk_ev[c], 0) != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "stream ordering failed");
        const uint64_t n_cells = L.capacity;
        const unsigned aux_grid = (unsigned)std::min<uint64_t>((n_cells + 255) / 256, (uint64_t)es.ctx->sm_count * 8);
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
                d_tb = (uint32_t*)es.cache_alloc((nseg + 1) * 4);
                if (!d_tb) return -tagg_fail(TAGG_ERR_OOM, "tile table allocation failed");
                if (cudaMemcpyAsync(d_tb, es.pin(tb.data(), (nseg + 1) * 4), (nseg + 1) * 4, cudaMemcpyHostToDevice, es.st) != cudaSuccess)
                    return -tagg_fail(TAGG_ERR_CUDA, "tile table upload failed");
                MParams sp = p;
                sp.segs = es.d_segs;
                sp.tile_begin = d_tb;
                sp.n_segs = (uint32_t)nseg;
                sp.n_tiles = (uint32_t)tiles;
                const uint64_t units = (tiles + n_sub - 1) / n_sub;
                const uint64_t sms = (uint64_t)es.ctx->sm_count > es.reserve_sms ? (uint64_t)es.ctx->sm_count - es.reserve_sms : 1;
                const uint32_t grid = (uint32_t)std::min<uint64_t>(sms * per_sm, units);
                fn<<<grid, threads, smem_bytes, es.st>>>(sp);
                cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "k_mterms launch failed: %s", cudaGetErrorString(e));
                es.ctx->launches++;
                es.n_launches++;
            }
        }
        {
            MFix fx;
            memset(&fx, 0, sizeof(fx));
            fx.n = n_cells;
            if (p.present_mode == PRESENT_COUNTS) { fx.present = es.arena + L.off_present; fx.counts = p.count_acc[0]; }
            for (int g = 0; g < p.n_groups; g++) {
                const MGroup& G = p.groups[g];
                if (G.seen_mode != SEEN_DERIVED) continue;
                auto& d = fx.g[fx.n_groups++];
                d.cell = G.derive_op == MO_SUM ? G.acc_sum : G.derive_op == MO_MIN ? G.acc_min : G.acc_max;
                d.seen = G.seen;
                d.derive_op = G.derive_op;
            }
            if (fx.counts || fx.n_groups) {
                k_mterms_fixup<<<aux_grid, 256, 0, es.st>>>(fx);
                cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "k_mterms_fixup launch failed: %s", cudaGetErrorString(e));
                es.ctx->launches++;
                es.n_launches++;
            }
        }
        es.skip[mem] = 1;
        handled++;
    }
    return handled;
}

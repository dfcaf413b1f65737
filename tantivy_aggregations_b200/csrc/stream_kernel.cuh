// stream_kernel.cuh — the streaming kernel template `k_stream<Shp<..>>` and its launch parameters (see stream.cu for the
// design notes).  Included by stream.cu (host side, shape selection) and by the stream_inst_*.cu translation units, each of
// which instantiates one family of shapes (the ~90 instantiations compile in parallel instead of in one 3-minute unit).
#pragma once
#include <type_traits>

#include "exec.h"

#define ST_WARPS 8                                   // consumer warps per group
#define ST_GROUP_THREADS ((ST_WARPS + 1) * 32)       // + 1 producer warp
#define ST_MAXGROUPS 3
#define ST_TILE TAGG_TILE_DOCS
#define ST_WORDS_PER_WARP (ST_TILE / 32 / ST_WARPS)  // 8
#define ST_DOCS_PER_WARP (ST_WORDS_PER_WARP * 32)    // 256
#define ST_MAXSTAGES 6
#define ST_MAXCOLS 6
#define ST_MAXPRED 4
#define ST_MAXBITS (2 + ST_MAXPRED)
#define ST_MAXRG 4
#define ST_MAXBG 3
#define ST_TBUF 128                                  // BK_RANK: out-of-range codes buffered per warp

enum { PR_FILTER = 0, PR_RANGE = 1, PR_LUT = 2 };
enum { OPB_SUM = 1, OPB_MIN = 2, OPB_MAX = 4 };
enum { BK_NONE = 0, BK_TERMS = 1, BK_HIST = 2, BK_RANK = 3 };
enum { SF_MAIN_BITS = 1, SF_DELETES = 2, SF_PRED_BITS0 = 4 /* << i */, SF_PRED_NONE0 = 256 /* << i */,
       SF_FPOS = 4096 /* every staged f64 column of the segment lies in [+0.0, +inf]: no sign handling, no NaN */ };

// Everything the kernel needs to know about one segment, prepared on the host.
struct SegDesc {
    uint32_t tile_begin, max_doc, flags, tile_bytes;  // tile_bytes: bytes one staged tile of this segment moves
    const uint8_t* col_ptr[ST_MAXCOLS];
    uint64_t minv[ST_MAXCOLS];
    uint32_t nb[ST_MAXCOLS];
    const uint8_t* bits_ptr[ST_MAXBITS];  // 0 = main docset, 1 = deleted, 2+i = filter docset of pred i
    // bytes readable behind bits_ptr (a multiple of 16): whole tiles for device bitsets; ceil16(max_doc / 8) for a
    // page-locked HOST bitset that the producer reads in place over PCIe (no staging copy, exec.cu normalise_docset)
    uint32_t bits_len[ST_MAXBITS];
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
    uint32_t tile_base;  // tile_begin of the first segment of this launch (chunked execute)
    int32_t n_cols;
    uint32_t soff_col[ST_MAXCOLS];  // byte offset of each staged column inside a stage
    uint32_t soff_bits;             // bitset slots (256 B each) inside a stage
    uint32_t stage_bytes;
    uint32_t n_stages, group_bytes, table_bytes;   // shared-memory layout
    uint32_t soff_tab_count[2];                    // STAB: CTA-private bucket count tables (u32)
    uint32_t soff_tab_sum[ST_MAXBG];               // STAB: CTA-private bucket sum tables (u64 / f64)
    uint32_t soff_tab_min[ST_MAXBG], soff_tab_max[ST_MAXBG];  // STAB: CTA-private min / max tables (u64, max-form)
    // STAB with filter tables: the shared min / max tables hold only a 32-bit rank word of the CTA's best value (u32) and
    // act as a filter in front of the exact global cells (a value whose high word does not reach the filter cannot be a
    // new extreme); survivors — a few per bucket and CTA — go to the global table with a checked atomic.  Halves the
    // table footprint so that a third consumer group fits (C2: 0.253 -> 0.21 ms)
    // The filter word is the value's position inside the column's code range [filt_lo, filt_hi] over all segments of the
    // call, scaled to 32 bits: (code - lo) >> shift for max, (hi - code) >> shift for min (so integer columns filter too)
    uint32_t tab_filt;
    uint32_t filt_shift[ST_MAXBG];
    uint64_t filt_lo[ST_MAXBG], filt_hi[ST_MAXBG];
    int32_t n_preds, n_vpreds;                     // all predicates / those evaluated on column values
    int32_t pred_type[ST_MAXPRED];
    int32_t pred_scol[ST_MAXPRED];
    const uint8_t* pred_lut[ST_MAXPRED];
    int32_t n_root_counts;
    uint64_t* root_count_acc[2];
    SGroup rgroups[ST_MAXRG];
    int32_t key_scol;
    uint64_t dom_min, dom_size;
    double f0, f1;
    // BK_RANK (percentiles, pct.cu): monotone code bins umulhi((code - rank_lo) >> rank_shift, rank_mul) over [rank_lo, rank_lo + rank_span);
    // codes outside are appended to an exact list (tail_count[0] = appended, [1] = those below rank_lo)
    uint64_t rank_lo, rank_span;
    uint32_t rank_shift, rank_mul;
    uint32_t rank_linear;   // bins equal-width in the value instead: (uint32)((f64(code) - rank_flo) * rank_fscale), clamped
    double rank_flo, rank_fscale;
    uint64_t* tail_codes;
    unsigned long long* tail_count;
    uint64_t tail_cap;
    uint32_t* overflow_flag;
    // BK_HIST with few buckets: hist_bounds[j] = smallest code whose ordinal is >= dom_min + j (j = 0..dom_size; entry
    // dom_size closes the valid range) — the exact IEEE division of histogram.rs:146 becomes a multiply + table fix-up
    const uint64_t* hist_bounds;
    uint32_t soff_hist_bounds;
    double hist_inv;
    // BK_RANK + histogram_agg_f64(same column, count_agg()) of the same tuple, fused into the percentile pass:
    // bucket counts in a second shared table (boundary-table ordinals, side_dom buckets from ordinal side_dom_min)
    uint32_t side_dom, soff_side_count;
    uint64_t side_dom_min;
    uint64_t* side_count_acc;
    uint8_t* side_present;
    // global tables (too many buckets for shared tables) with min / max on bucket group 0: one byte per bucket in shared
    // memory holds the best 4-bit LEVEL seen by this CTA — level = (code - nib_lo) >> nib_shift clamped to 15, high nibble
    // for max, low nibble (15 - level) for min.  Only values whose level reaches the filter touch the global cells (a
    // fire-and-forget RED); updates of the byte are plain stores (a lost update only weakens the filter)
    uint32_t soff_nib, nib_shift;
    uint64_t nib_lo;
    // nib_hi32: the span is wide enough that the level can be taken from the high word alone, (code_hi - nib_lo_hi) >>
    // (nib_shift - 32) — any monotone function of the code is a valid filter level.  present_from_nib: both nibbles of a
    // touched bucket's byte are raised (level + (15 - level) == 15, never both zero), so the byte doubles as the bucket-
    // existence record and the CTA bitmap is not needed
    uint32_t nib_hi32, present_from_nib;
    uint32_t soff_present_bits;  // global tables without a count: CTA bitmap of touched buckets (0 = none), flushed at the end
    uint8_t* present;      // maintained by the kernel (global tables without a count); nullptr otherwise
    uint8_t* present_out;  // STAB: written by the final table merge
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
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar_saddr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_saddr, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar_saddr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// What the producer warp tells the consumers about a staged tile.
struct TileDesc {
    uint32_t n_valid, flags;
    uint32_t nb[ST_MAXCOLS];
    uint64_t minv[ST_MAXCOLS];
    uint64_t pred_lo[ST_MAXPRED], pred_hi[ST_MAXPRED];
    uint32_t seg, pad;  // segment of the tile: consumers rebuild their column descriptors only when it changes
};

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// A staged column of the current tile, held in registers (32-bit shared-memory address, split masks).
struct TCol {
    uint32_t saddr, nb, mlo, mhi, minlo, minhi;
};
// TileDesc fields are read with explicit shared-memory loads (a generic-pointer load is tracked on the
// long scoreboard and costs a global-memory-class latency)
#define TD_N_VALID 0
#define TD_FLAGS 4
#define TD_NB(c) (8 + 4 * (c))
#define TD_MINV(c) (8 + 4 * ST_MAXCOLS + 8 * (c))
#define TD_PRED_LO(i) (8 + 12 * ST_MAXCOLS + 8 * (i))
#define TD_PRED_HI(i) (8 + 12 * ST_MAXCOLS + 8 * ST_MAXPRED + 8 * (i))
#define TD_SEG (8 + 12 * ST_MAXCOLS + 16 * ST_MAXPRED)
static_assert(sizeof(TileDesc) == 16 + 12 * ST_MAXCOLS + 16 * ST_MAXPRED, "TileDesc layout");
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ TCol tcol(const SParams& p, uint32_t T, uint32_t stage_saddr, int scol) {
    TCol c;
    c.saddr = stage_saddr + p.soff_col[scol];
    c.nb = lds32(T + TD_NB(scol));
    uint64_t mn = lds64(T + TD_MINV(scol));
    c.minlo = (uint32_t)mn;
    c.minhi = (uint32_t)(mn >> 32);
    c.mlo = c.nb >= 32 ? 0xffffffffu : ((1u << c.nb) - 1u);
    c.mhi = c.nb <= 32 ? 0u : (c.nb >= 64 ? 0xffffffffu : ((1u << (c.nb - 32)) - 1u));
    return c;
}
// packed delta of value i (tantivy BitUnpacker::get on shared memory, 32-bit aligned loads)
__device__ __forceinline__ void tdelta(const TCol& c, uint32_t i, uint32_t& lo, uint32_t& hi) {
    uint32_t bit = i * c.nb;
    uint32_t a = c.saddr + ((bit >> 5) << 2), sh = bit & 31u;
    uint32_t w0 = lds32(a), w1 = lds32(a + 4);
    lo = __funnelshift_r(w0, w1, sh) & c.mlo;
    hi = 0;
    if (c.nb > 32) {
        uint32_t w2 = lds32(a + 8);
        hi = __funnelshift_r(w1, w2, sh) & c.mhi;
    }
}
__device__ __forceinline__ uint64_t tget(const TCol& c, uint32_t i) {  // -> code
    uint32_t lo, hi;
    tdelta(c, i, lo, hi);
    return (((uint64_t)hi << 32) | lo) + (((uint64_t)c.minhi << 32) | c.minlo);
}

// Kernel shape.  RT shapes read the per-group op masks from the launch parameters (any flat plan);
// CT shapes bake the op masks of the single bucket / root column group into the instantiation, so the
// compiler drops every path the plan does not have (the hot configurations use these).
template <int BUCKET_, int NBG_, int NRG_, bool COMPACT_, bool STAB_, int BOPS_ = -1, int ROPS_ = -1, int FILT_ = -1>
struct Shp {
    static constexpr int BUCKET = BUCKET_, NBG = NBG_, NRG = NRG_;
    static constexpr bool COMPACT = COMPACT_, STAB = STAB_;
    static constexpr int BOPS = BOPS_, ROPS = ROPS_;
    static constexpr int FILT = FILT_;  // filter tables: -1 decided by the launch parameters (RT shapes), 0 / 1 compiled in (CT shapes)
};

// Warp-specialised: a CTA is G groups of 9 warps — warp 0 of a group is the TMA producer, warps 1..8
// consume.  Stages are handed over with mbarriers only (full: TMA bytes landed; empty: 8 consumer
// warps released the stage), so consumer warps never synchronise with one another inside the loop.
// STAB: the bucket tables (counts u32; sums / min / max u64) live in shared memory, private to the CTA,
// and are merged into the global tables once at the end (global atomics on a few thousand hot addresses
// serialise in L2 and L2 reads of them cap near 130 G/s; shared-memory atomics do not —
// tools/atom_bench.cu).
template <class SH>
__global__ void __launch_bounds__(ST_GROUP_THREADS * ST_MAXGROUPS) k_stream(const __grid_constant__ SParams p) {
    constexpr int BUCKET = SH::BUCKET, NBG = SH::NBG, NRG = SH::NRG;
    constexpr bool COMPACT = SH::COMPACT, STAB = SH::STAB;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint32_t n_groups = blockDim.x / ST_GROUP_THREADS;
    const uint32_t group = tid / ST_GROUP_THREADS;
    const uint32_t gwarp = (tid % ST_GROUP_THREADS) >> 5;  // 0 = producer, 1..8 = consumers
    const uint32_t S = p.n_stages;

    // shared layout: [tables][per group: stages | queues | tile descs | barriers]
    uint8_t* gbase = smem + p.table_bytes + (size_t)group * p.group_bytes;
    uint8_t* stages = gbase;
    uint16_t* queues = (uint16_t*)(gbase + (size_t)S * p.stage_bytes);
    TileDesc* tdesc = (TileDesc*)(queues + ST_WARPS * ST_DOCS_PER_WARP);
    uint64_t* full = (uint64_t*)(tdesc + S);
    uint64_t* empty = full + S;

    if (tid % ST_GROUP_THREADS == 0) {
        for (uint32_t s = 0; s < S; s++) { mbar_init(full + s, 1); mbar_init(empty + s, ST_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // shared tables (STAB) / presence bitmap / histogram boundaries live in front of the group blocks
        uint32_t* t32 = (uint32_t*)smem;
        for (uint32_t i = tid; i < p.table_bytes / 4; i += blockDim.x) t32[i] = 0;
        if ((BUCKET == BK_HIST || BUCKET == BK_RANK) && p.hist_bounds) {
            __syncthreads();
            uint64_t* b = (uint64_t*)(smem + p.soff_hist_bounds);
            const uint32_t nb = BUCKET == BK_HIST ? (uint32_t)p.dom_size : p.side_dom;
            for (uint32_t i = tid; i <= nb; i += blockDim.x) b[i] = p.hist_bounds[i];
        }
        if (STAB && NBG > 0) {  // f64 sums fold from -0.0 (dev.cuh F64_NEG_ZERO_BITS)
            bool any = false;
#pragma unroll
            for (int g = 0; g < NBG; g++) {
                const uint32_t ops = (g == 0 && SH::BOPS >= 0) ? (uint32_t)SH::BOPS : p.bgroups[g].ops;
                if ((ops & OPB_SUM) && p.bgroups[g].kind == TAGG_F64) {
                    if (!any) __syncthreads();
                    any = true;
                    uint64_t* t = (uint64_t*)(smem + p.soff_tab_sum[g]);
                    for (uint32_t i = tid; i < (uint32_t)p.dom_size; i += blockDim.x) t[i] = F64_NEG_ZERO_BITS;
                }
            }
        }
    }
    __syncthreads();

    const uint64_t first = (uint64_t)blockIdx.x * n_groups + group, step = (uint64_t)gridDim.x * n_groups;

    if (gwarp == 0) {
        // ======================= producer warp =======================
        // A single thread preparing a tile is a ~1000-cycle serial chain, which caps a group at one tile
        // per microsecond.  The work is spread over the lanes instead: lane c issues the TMA copy of
        // staged column c, lane 8+b that of bitset b, lanes 16.. publish the tile descriptor, lane 0 arms
        // the barrier.  (A copy may complete before the barrier is armed: the phase cannot flip while the
        // producer's own arrival is pending.)
        uint32_t cur_seg = 0, stage = 0, parity = 1;  // parity of the empty-barrier phase to wait for (first lap: none)
        bool first_lap = true;
        for (uint64_t tile = first; tile < p.n_tiles; tile += step) {
            if (!first_lap) mbar_wait(empty + stage, parity);
            const uint32_t gt = (uint32_t)tile + p.tile_base;  // tile index over the whole call
            while (cur_seg + 1 < p.n_segs && p.segs[cur_seg + 1].tile_begin <= gt) cur_seg++;
            const SegDesc* Sg = p.segs + cur_seg;
            const uint32_t lt = gt - Sg->tile_begin;
            const uint32_t flags = Sg->flags;
            TileDesc* T = tdesc + stage;
            uint8_t* base = stages + (size_t)stage * p.stage_bytes;
            uint32_t issued = 0;  // bytes this lane asked the TMA unit for
            if (lane < (uint32_t)p.n_cols) {
                uint32_t cb = (ST_TILE / 8) * Sg->nb[lane];
                if (cb) tma_bulk_g2s(base + p.soff_col[lane], Sg->col_ptr[lane] + (size_t)lt * cb, cb, full + stage);
                issued = cb;
            } else if (lane >= 8 && lane < 8 + ST_MAXBITS) {
                uint32_t b = lane - 8;
                if (flags & (1u << b)) {
                    const uint32_t at = lt * (ST_TILE / 8), len = Sg->bits_len[b];
                    const uint32_t cb = len > at ? min((uint32_t)(ST_TILE / 8), len - at) : 0u;  // a host bitset ends inside its last tile
                    if (cb) tma_bulk_g2s(base + p.soff_bits + 256 * b, Sg->bits_ptr[b] + at, cb, full + stage);
                    issued = cb;
                }
            } else if (lane == 16) {
                T->n_valid = (uint32_t)min((uint64_t)ST_TILE, (uint64_t)Sg->max_doc - (uint64_t)lt * ST_TILE);
                T->flags = flags;
                T->seg = cur_seg;
            } else if (lane >= 17 && lane < 17 + ST_MAXCOLS) {
                uint32_t c = lane - 17;
                if (c < (uint32_t)p.n_cols) { T->nb[c] = Sg->nb[c]; T->minv[c] = Sg->minv[c]; }
            } else if (lane >= 24 && lane < 24 + ST_MAXPRED) {
                uint32_t i = lane - 24;
                if (i < (uint32_t)p.n_preds) { T->pred_lo[i] = Sg->pred_lo[i]; T->pred_hi[i] = Sg->pred_hi[i]; }
            }
            const uint32_t bytes = __reduce_add_sync(0xffffffffu, issued);
            if (lane == 0) {
                if (bytes) mbar_expect_tx(full + stage, bytes);
                else mbar_arrive(full + stage);  // nothing to stage (count over AllQuery)
            }
            if (++stage == S) { stage = 0; parity ^= 1u; first_lap = false; }
        }
    } else {
        // ======================= consumer warps =======================
        const uint32_t warp = gwarp - 1;
        const uint32_t lane_bit = 1u << lane, lt_mask = lane_bit - 1u;
        const uint32_t q_saddr = smem_u32(queues + warp * ST_DOCS_PER_WARP);
        const uint32_t stages_saddr = smem_u32(stages);
        const uint32_t tdesc_saddr = smem_u32(tdesc);
        const uint32_t smem_saddr = smem_u32(smem);
        // op masks: compile-time for CT shapes (single group), launch parameters otherwise
        const uint32_t ops_b0 = SH::BOPS >= 0 ? (uint32_t)SH::BOPS : (NBG > 0 ? p.bgroups[0].ops : 0u);
        const uint32_t ops_b1 = NBG > 1 ? p.bgroups[1].ops : 0u, ops_b2 = NBG > 2 ? p.bgroups[2].ops : 0u;
        const uint32_t ops_r0 = SH::ROPS >= 0 ? (uint32_t)SH::ROPS : (NRG > 0 ? p.rgroups[0].ops : 0u);
        const uint32_t dom_size32 = (uint32_t)p.dom_size;
        const bool filt = SH::FILT < 0 ? (STAB && p.tab_filt != 0) : (SH::FILT == 1);
        const bool two_counts = SH::BOPS < 0 && p.n_bcounts > 1;  // CT shapes are picked for at most one bucket count

        // per-thread root accumulators
        uint64_t rsum[NRG ? NRG : 1], rmin[NRG ? NRG : 1], rmax[NRG ? NRG : 1];
        bool rseen = false;
#pragma unroll
        for (int g = 0; g < NRG; g++) { rsum[g] = p.rgroups[g].kind == TAGG_F64 ? F64_NEG_ZERO_BITS : 0ull; rmin[g] = 0; rmax[g] = 0; }
        // CT root shape (one f64 column, compile-time ops): min / max run on the packed deltas and are folded into
        // the code domain only when the column's min_value changes (segment change); the sum adds delta + constant
        constexpr bool CTROOT = SH::ROPS >= 0 && NRG == 1;
        constexpr bool RANK_LINEAR = BUCKET == BK_RANK && SH::ROPS == -2;  // value-space rank bins (pct.cu)
        uint64_t dmin = ~0ull, dmax = 0, cminv = 0;
        bool fseen = false;
        auto fold_root = [&]() {
            if (fseen) {
                const uint64_t cmax = dmax + cminv, cmin = ~(dmin + cminv);
                rmax[0] = cmax > rmax[0] ? cmax : rmax[0];
                rmin[0] = cmin > rmin[0] ? cmin : rmin[0];
            }
            dmin = ~0ull; dmax = 0; fseen = false;
        };
        bool bad_key = false;  // a key outside the domain its column header declares (corrupt input)
        uint32_t matched = 0;  // every lane holds the warp's count

        // BK_RANK: this warp's buffer of out-of-range codes (the last ST_WARPS * ST_TBUF * 8 bytes of the group block)
        const uint32_t tbuf_saddr = smem_u32(gbase + p.group_bytes - (ST_WARPS - warp) * ST_TBUF * 8);
        uint32_t wtail = 0;
        TCol c_kc = {0, 0, 0, 0, 0, 0}, c_bc[NBG ? NBG : 1], c_rc[NRG ? NRG : 1];
        uint32_t c_seg = 0xffffffffu, c_base = 0, c_krel = 0;
        uint32_t c_ptab = 0;  // cached truth tables of the bit-plane predicates (see the mask phase)
#pragma unroll
        for (int g = 0; g < (NBG ? NBG : 1); g++) c_bc[g] = c_kc;
#pragma unroll
        for (int g = 0; g < (NRG ? NRG : 1); g++) c_rc[g] = c_kc;
        const uint32_t full_saddr = smem_u32(full), empty_saddr = smem_u32(empty);
        auto flush_tail = [&]() {  // warp-uniform: wtail buffered codes -> the global list
            __syncwarp();
            unsigned long long base = 0;
            uint32_t nlow = 0;
            for (uint32_t i = lane; i < wtail; i += 32) nlow += lds64(tbuf_saddr + 8 * i) < p.rank_lo ? 1u : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nlow += __shfl_xor_sync(0xffffffffu, nlow, o);
            if (lane == 0) {
                base = atomicAdd(p.tail_count, (unsigned long long)wtail);
                if (nlow) atomicAdd(p.tail_count + 1, (unsigned long long)nlow);
            }
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t i = lane; i < wtail; i += 32) {
                if (base + i < p.tail_cap) p.tail_codes[base + i] = lds64(tbuf_saddr + 8 * i);
                else *p.overflow_flag = 3u;
            }
            __syncwarp();
            wtail = 0;
        };
        uint32_t stage = 0, parity = 0;
        for (uint64_t tile = first; tile < p.n_tiles; tile += step) {
            mbar_wait_s(full_saddr + 8 * stage, parity);
            const uint32_t stage_saddr = stages_saddr + stage * p.stage_bytes;
            const uint32_t T = tdesc_saddr + stage * (uint32_t)sizeof(TileDesc);
            const uint32_t flags = lds32(T + TD_FLAGS), n_valid = lds32(T + TD_N_VALID);
            const uint32_t bits_saddr = stage_saddr + p.soff_bits;

            // ---- phase 1: one match-mask word per lane (lanes 0..7) --------------------------------
            uint32_t m = 0;
            if ((COMPACT || n_valid != ST_TILE || p.n_vpreds) && lane < ST_WORDS_PER_WARP) {
                const uint32_t wi = warp * ST_WORDS_PER_WARP + lane;
                const uint32_t d0 = wi * 32;
                const uint32_t wa = bits_saddr + wi * 4;
                m = d0 + 32 <= n_valid ? 0xffffffffu : (d0 >= n_valid ? 0u : ((1u << (n_valid - d0)) - 1u));
                if (flags & SF_MAIN_BITS) m &= lds32(wa);
                if (flags & SF_DELETES) m &= ~lds32(wa + 256);  // searcher.rs:41-46
                if (flags & (SF_PRED_BITS0 << 0)) m &= lds32(wa + 512);
                if (flags & (SF_PRED_BITS0 << 1)) m &= lds32(wa + 768);
                if (flags & (SF_PRED_BITS0 << 2)) m &= lds32(wa + 1024);
                if (flags & (SF_PRED_BITS0 << 3)) m &= lds32(wa + 1280);
                if (flags & (SF_PRED_NONE0 * 15u)) m = 0;  // a filter query that matches nothing in this segment
            }
            // value predicates (post_filter / COLUMN_RANGE docsets): per document, folded in with ballots
            const uint32_t tseg = lds32(T + TD_SEG);
            if (p.n_vpreds) {
                // bit planes of a 1- or 2-bit column under a truth table `tt` over its values -> mask word of documents
                // [32 lane, 32 lane + 32) for lanes < 8
                auto planes = [&](uint32_t nb, uint32_t tt, uint32_t saddr) -> uint32_t {
                    uint32_t pm = 0;
                    if (lane < ST_WORDS_PER_WARP) {
                        if (nb == 1) {
                            const uint32_t x = lds32(saddr + (warp * ST_WORDS_PER_WARP + lane) * 4);
                            pm = ((tt & 1u) ? ~x : 0u) | ((tt & 2u) ? x : 0u);
                        } else {
                            const uint64_t x = lds64(saddr + (warp * ST_WORDS_PER_WARP + lane) * 8);
                            const uint32_t M0 = (tt & 1u) ? 0x55555555u : 0u, M1 = (tt & 2u) ? 0x55555555u : 0u;
                            const uint32_t M2 = (tt & 4u) ? 0x55555555u : 0u, M3 = (tt & 8u) ? 0x55555555u : 0u;
                            auto half = [&](uint32_t h) {
                                const uint32_t b0 = h, b1 = h >> 1;  // (the table masks keep the even bit positions only)
                                uint32_t r = (M0 & ~b1 & ~b0) | (M1 & ~b1 & b0) | (M2 & b1 & ~b0) | (M3 & b1 & b0);
                                r = (r | (r >> 1)) & 0x33333333u;
                                r = (r | (r >> 2)) & 0x0f0f0f0fu;
                                r = (r | (r >> 4)) & 0x00ff00ffu;
                                r = (r | (r >> 8)) & 0x0000ffffu;
                                return r;
                            };
                            pm = half((uint32_t)x) | (half((uint32_t)(x >> 32)) << 16);
                        }
                    }
                    return pm;
                };
                for (int i = 0; i < p.n_preds; i++) {
                    const int type = p.pred_type[i];
                    if (type == PR_FILTER) continue;
                    // a truth table depends on the segment only (column min_value, predicate bounds): tiles of the same segment
                    // reuse it — byte i of c_ptab = 0x80 | width << 4 | table
                    // (not in the rank-bin kernels: they are instruction-cache sensitive, +5 % with this path compiled in)
                    const uint32_t pb = (c_ptab >> (8 * i)) & 0xffu;
                    if (BUCKET != BK_RANK && tseg == c_seg && (pb & 0x80u)) {
                        m &= planes((pb >> 4) & 3u, pb & 15u, stage_saddr + p.soff_col[p.pred_scol[i]]);
                        continue;
                    }
                    if (BUCKET != BK_RANK) c_ptab &= ~(0xffu << (8 * i));
                    const TCol pc = tcol(p, T, stage_saddr, p.pred_scol[i]);
                    const uint64_t lo = lds64(T + TD_PRED_LO(i)), hi = lds64(T + TD_PRED_HI(i));
                    const uint8_t* lut = p.pred_lut[i];
                    if (pc.nb >= 1 && pc.nb <= 2) {
                        // One- and two-bit columns (status-like fields of <= 4 values): the predicate is a truth table over the
                        // column's values, evaluated once per tile (lane v tests value v), and applied to 32 documents per lane
                        // at once on the bit planes of the packed stream — no per-document work at all.  Lane j < 8 produces
                        // the mask word of documents [32j, 32j + 32) directly in the word-per-lane layout.
                        const uint64_t minv = ((uint64_t)pc.minhi << 32) | pc.minlo;
                        bool tv = false;
                        if (lane < (1u << pc.nb)) {
                            const uint64_t code = minv + lane;
                            if (type == PR_LUT) {
                                const uint64_t r = code - lo;
                                tv = code >= lo && r < hi && ((lut[r >> 3] >> (r & 7)) & 1);
                            } else {
                                tv = code >= lo && code <= hi;
                            }
                        }
                        const uint32_t tt = __ballot_sync(0xffffffffu, tv) & 15u;
                        if (BUCKET != BK_RANK) c_ptab |= (0x80u | (pc.nb << 4) | tt) << (8 * i);
                        m &= planes(pc.nb, tt, pc.saddr);
                        continue;
                    }
                    if (pc.nb >= 1 && pc.nb <= 8 && (type == PR_RANGE || hi <= 32)) {
                        // Narrow column (status-like fields): lane l tests documents [8l, 8l + 8) of the warp's 256 from one
                        // 64-bit window of the packed stream, on the packed deltas (the predicate's code range is moved
                        // into the delta domain once per tile; a LUT of <= 32 entries lives in a register), then the
                        // result bytes are regrouped into the word-per-lane layout.  ~4x fewer instructions than a
                        // ballot per 32 documents.
                        const uint64_t minv = ((uint64_t)pc.minhi << 32) | pc.minlo;
                        const uint64_t last = type == PR_LUT ? lo + hi - 1 : hi;  // last code that can pass (LUT: hi = entries > 0 here)
                        uint32_t vlo = 1, vspan = 0, roff = 0, lutw = 0xffffffffu;  // empty unless set below
                        bool any = type == PR_LUT ? hi > 0 && last >= lo : hi >= lo;
                        any = any && last >= minv && lo <= minv + 255;
                        if (any) {
                            vlo = lo > minv ? (uint32_t)(lo - minv) : 0u;
                            const uint32_t vhi = last - minv > 255 ? 255u : (uint32_t)(last - minv);
                            vspan = vhi - vlo;
                            if (type == PR_LUT) {
                                roff = (uint32_t)(minv + vlo - lo);  // LUT index of delta vlo
                                lutw = (uint32_t)lut[0] | ((uint32_t)lut[1] << 8) | ((uint32_t)lut[2] << 16) | ((uint32_t)lut[3] << 24);
                            }
                        }
                        const uint32_t B = (warp * 32 + lane) * pc.nb;  // byte offset of this lane's 8 values
                        const uint32_t a = pc.saddr + (B & ~3u), sh = (B & 3u) * 8;
                        const uint32_t w0 = lds32(a), w1 = lds32(a + 4), w2 = lds32(a + 8);
                        const uint64_t win = ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
                        uint32_t okb = 0;
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const uint32_t t = ((uint32_t)(win >> (k * pc.nb)) & pc.mlo) - vlo;  // delta - vlo, wraps below
                            const bool ok = any && t <= vspan && ((lutw >> ((t + roff) & 31u)) & 1u);
                            okb |= ok ? (1u << k) : 0u;
                        }
                        // word j of the warp = bytes of lanes 4j .. 4j + 3
                        uint32_t pm = 0;
#pragma unroll
                        for (int t = 0; t < 4; t++) pm |= __shfl_sync(0xffffffffu, okb, (lane * 4 + t) & 31) << (8 * t);
                        m &= pm;
                        continue;
                    }
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
            }

            // column descriptors of the roles this instantiation has, in registers; they depend on the segment only
            // (bit width, min_value), so a tile of the same segment just moves the shared-memory address
            TCol kc, bc[NBG ? NBG : 1], rc[NRG ? NRG : 1];
            uint32_t krel = 0;  // TERMS: (column min - domain min); keys are dense and < 2^24 wide
            if (tseg != c_seg) {
                if (BUCKET != BK_NONE) {
                    kc = tcol(p, T, stage_saddr, p.key_scol);
                    krel = (uint32_t)(lds64(T + TD_MINV(p.key_scol)) - p.dom_min);
                }
#pragma unroll
                for (int g = 0; g < NBG; g++) bc[g] = tcol(p, T, stage_saddr, p.bgroups[g].scol);
#pragma unroll
                for (int g = 0; g < NRG; g++) rc[g] = tcol(p, T, stage_saddr, p.rgroups[g].scol);
                c_seg = tseg; c_base = stage_saddr; c_kc = kc; c_krel = krel;
#pragma unroll
                for (int g = 0; g < NBG; g++) c_bc[g] = bc[g];
#pragma unroll
                for (int g = 0; g < NRG; g++) c_rc[g] = rc[g];
            } else {
                const uint32_t moved = stage_saddr - c_base;
                kc = c_kc; kc.saddr += moved; krel = c_krel;
#pragma unroll
                for (int g = 0; g < NBG; g++) { bc[g] = c_bc[g]; bc[g].saddr += moved; }
#pragma unroll
                for (int g = 0; g < NRG; g++) { rc[g] = c_rc[g]; rc[g].saddr += moved; }
            }

            // every f64 column this tile touches lies in [+0.0, +inf] (host-checked against the column headers):
            // code -> f64 is one XOR, and f64 min / max agree with the order of the codes
            const bool fpos = (flags & SF_FPOS) != 0;
            uint64_t rbase = 0;  // CT root: f64 bits = delta + rbase when fpos
            if (CTROOT) {
                const uint64_t mv = ((uint64_t)rc[0].minhi << 32) | rc[0].minlo;
                if (mv != cminv) { fold_root(); cminv = mv; }
                rbase = mv ^ 0x8000000000000000ull;
            }
            // histogram: ordinal via multiply + boundary fix-up (exact), see SParams::hist_bounds
            const bool hb = (BUCKET == BK_HIST || BUCKET == BK_RANK) && p.hist_bounds != nullptr;
            const uint32_t hb_saddr = smem_saddr + p.soff_hist_bounds;
            const uint32_t hb_dom = BUCKET == BK_HIST ? dom_size32 : p.side_dom;
            uint64_t hb_first = 0, hb_end = 0;
            if (hb) { hb_first = lds64(hb_saddr); hb_end = lds64(hb_saddr + 8 * hb_dom); }
            const double hb_dmin = (double)(BUCKET == BK_HIST ? p.dom_min : p.side_dom_min);

            // U matched documents per lane at a time (independent chains overlap the table latency).
            // CHECK: slots may be empty (act[u] false) — the ragged tail; otherwise every slot is live.
            auto process = [&](auto U_, auto CHECK_, auto POS_, const uint32_t* dl, const bool* act_in) {
                constexpr int U = decltype(U_)::value;
                constexpr bool CHECK = decltype(CHECK_)::value;
                constexpr bool POS = decltype(POS_)::value;
                // (BK_RANK keeps ONE instantiation and branches on the per-tile flag: its code already crowds the instruction cache)
                auto c2f = [&](uint64_t code) {
                    if (BUCKET == BK_RANK) return fpos ? __longlong_as_double((long long)(code ^ 0x8000000000000000ull)) : code_to_f64(code);
                    return POS ? __longlong_as_double((long long)(code ^ 0x8000000000000000ull)) : code_to_f64(code);
                };
                // histogram ordinal of a code relative to the first bucket, -1: skipped (NaN or below start, histogram.rs:138-145).
                // A multiply lands next to the exact ordinal; the boundary table (exact, monotone in the code) decides.
                auto hist_bin = [&](uint64_t code) -> int {
                    const double t = __dsub_rn(__dmul_rn(__dsub_rn(c2f(code), p.f0), p.hist_inv), hb_dmin);
                    uint32_t r = (uint32_t)min(max(__double2int_rz(t), 0), (int)hb_dom - 1);
                    const uint64_t b0 = lds64(hb_saddr + 8 * r), b1 = lds64(hb_saddr + 8 * r + 8);
                    if (code < b0 || code >= b1) {
                        if (code < hb_first || code >= hb_end) return -1;
                        while (code < lds64(hb_saddr + 8 * r)) r--;
                        while (code >= lds64(hb_saddr + 8 * (r + 1))) r++;
                    }
                    return (int)r;
                };
                bool act[U];
#pragma unroll
                for (int u = 0; u < U; u++) { act[u] = CHECK ? act_in[u] : true; rseen = rseen || act[u]; }
                if (CTROOT) {
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        if (act[u]) {
                            uint32_t lo, hi;
                            tdelta(rc[0], dl[u], lo, hi);
                            const uint64_t d = ((uint64_t)hi << 32) | lo;
                            if (SH::ROPS & OPB_SUM) {
                                const double v = POS ? __longlong_as_double((long long)(d + rbase)) : code_to_f64(d + cminv);
                                rsum[0] = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)rsum[0]), v));
                            }
                            if (SH::ROPS & OPB_MIN) dmin = d < dmin ? d : dmin;
                            if (SH::ROPS & OPB_MAX) dmax = d > dmax ? d : dmax;
                            fseen = true;
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < (CTROOT ? 0 : NRG); g++) {
                    const SGroup& G = p.rgroups[g];
                    const uint32_t ops = g == 0 ? ops_r0 : G.ops;
                    if (ops) {
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            if (act[u]) {
                                uint64_t code = tget(rc[g], dl[u]);
                                if (ops & OPB_SUM) {
                                    if (G.kind == TAGG_F64) rsum[g] = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)rsum[g]), c2f(code)));
                                    else rsum[g] += code_to_bits(G.kind, code);
                                }
                                if (ops & OPB_MIN) { uint64_t v = ~code; rmin[g] = v > rmin[g] ? v : rmin[g]; }
                                if (ops & OPB_MAX) rmax[g] = code > rmax[g] ? code : rmax[g];
                            }
                        }
                    }
                }
                if (BUCKET != BK_NONE) {
                    uint32_t rel[U];
                    bool tail[U];
                    uint64_t tail_code[U];
                    uint32_t rank_x[U];  // BK_RANK: 32-bit rank word of the code inside the binned range (doubles as the min / max filter word)
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        rel[u] = 0;
                        tail[u] = false;
                        tail_code[u] = 0;
                        rank_x[u] = 0;
                        if (act[u]) {
                            if (BUCKET == BK_TERMS) {
                                uint32_t lo, hi;
                                tdelta(kc, dl[u], lo, hi);
                                rel[u] = lo + krel;
                                // The key domain comes from the column headers, so a key cannot leave it.  Ragged batches keep
                                // the per-document guard; full batches stay branch-free (clamp + a sticky error flag that fails
                                // the call if a column ever contradicts its header)
                                if (CHECK) { act[u] = rel[u] < dom_size32; }
                                else { bad_key = bad_key || rel[u] >= dom_size32; rel[u] = min(rel[u], dom_size32 - 1u); }
                            } else if (BUCKET == BK_RANK) {
                                const uint64_t code = tget(kc, dl[u]);
                                const uint64_t d = code - p.rank_lo;  // below rank_lo: wraps above the span
                                rank_x[u] = (uint32_t)(d >> p.rank_shift);
                                if (RANK_LINEAR) {  // (its own instantiation: the rank kernels are instruction-cache sensitive)
                                    rel[u] = min(__double2uint_rz(__dmul_rn(__dsub_rn(c2f(code), p.rank_flo), p.rank_fscale)), dom_size32 - 1u);
                                } else {
                                    rel[u] = p.rank_mul ? __umulhi(rank_x[u], p.rank_mul) : rank_x[u];
                                }
                                tail[u] = d >= p.rank_span;
                                tail_code[u] = code;
                                if (hb) {  // the fused histogram counts every matched value, binned or not
                                    const int r = hist_bin(code);
                                    if (r >= 0) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_saddr + p.soff_side_count + 4 * (uint32_t)r) : "memory");
                                }
                            } else if (hb) {
                                const int r = hist_bin(tget(kc, dl[u]));
                                if (r < 0) act[u] = false;
                                else rel[u] = (uint32_t)r;
                            } else {
                                uint64_t ord;
                                // NaN or below start: skipped (histogram.rs:138-145)
                                if (hist_ord(tget(kc, dl[u]), p.f0, p.f1, &ord) && ord >= p.dom_min && ord - p.dom_min < p.dom_size) rel[u] = (uint32_t)(ord - p.dom_min);
                                else act[u] = false;
                            }
                        }
                    }
                    if (BUCKET == BK_RANK) {
                        // values outside the binned range go to the exact list, through a per-warp buffer in shared
                        // memory: one global atomic per ST_TBUF values (a single hot counter serialises in L2)
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            const uint32_t tm = __ballot_sync(0xffffffffu, tail[u]);
                            if (tm) {
                                if (wtail + __popc(tm) > ST_TBUF) flush_tail();
                                if (tail[u]) {
                                    const uint32_t at = wtail + __popc(tm & lt_mask);
                                    asm volatile("st.shared.u64 [%0], %1;" ::"r"(tbuf_saddr + 8 * at), "l"(tail_code[u]) : "memory");
                                    act[u] = false;
                                }
                                wtail += __popc(tm);
                            }
                        }
                    }
                    // bucket counts (STAB: also the bucket-existence record)
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        if (act[u]) {
                            if (STAB) {
                                asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_saddr + p.soff_tab_count[0] + 4 * rel[u]) : "memory");
                                if (two_counts) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(smem_saddr + p.soff_tab_count[1] + 4 * rel[u]) : "memory");
                            } else {
                                if (p.n_bcounts > 0) atomicAdd((unsigned long long*)(p.bcount_acc[0] + rel[u]), 1ull);
                                if (two_counts) atomicAdd((unsigned long long*)(p.bcount_acc[1] + rel[u]), 1ull);
                                if (p.soff_present_bits) {  // no count names the bucket: CTA bitmap, flushed once at the end
                                    const uint32_t wa = smem_saddr + p.soff_present_bits + 4 * (rel[u] >> 5), bit = 1u << (rel[u] & 31);
                                    if (!(lds32(wa) & bit)) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(wa), "r"(bit) : "memory");
                                } else if (!p.present_from_nib && p.present && !p.present[rel[u]]) {
                                    p.present[rel[u]] = 1;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int g = 0; g < NBG; g++) {
                        const SGroup& G = p.bgroups[g];
                        const uint32_t ops = g == 0 ? ops_b0 : g == 1 ? ops_b1 : ops_b2;
                        uint64_t code[U], cur_min[U], cur_max[U];
                        const bool nibf = !STAB && g == 0 && p.soff_nib != 0;
                        uint32_t nfb[U], nq[U];
                        // min / max cells: read all U of them first.  STAB: the CTA's shared table.  Otherwise
                        // global: a plain (L1, possibly stale) read filters most documents, survivors are
                        // confirmed at L2 before the atomic.
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            code[u] = 0; cur_min[u] = ~0ull; cur_max[u] = ~0ull; nfb[u] = 0; nq[u] = 0;
                            if (act[u]) {
                                code[u] = BUCKET == BK_RANK ? tail_code[u] : tget(bc[g], dl[u]);
                                if (nibf) {
                                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(nfb[u]) : "r"(smem_saddr + p.soff_nib + rel[u]));
                                    if (p.nib_hi32) {
                                        nq[u] = min(((uint32_t)(code[u] >> 32) - (uint32_t)(p.nib_lo >> 32)) >> (p.nib_shift - 32), 15u);
                                    } else {
                                        const uint64_t lv = code[u] >= p.nib_lo ? (code[u] - p.nib_lo) >> p.nib_shift : 0ull;
                                        nq[u] = lv > 15 ? 15u : (uint32_t)lv;
                                    }
                                } else if (filt) {  // cur_* hold the filter word; code[u] is compared through its own rank word below
                                    if (ops & OPB_MIN) cur_min[u] = lds32(smem_saddr + p.soff_tab_min[g] + 4 * rel[u]);
                                    if (ops & OPB_MAX) cur_max[u] = lds32(smem_saddr + p.soff_tab_max[g] + 4 * rel[u]);
                                } else {
                                    if (ops & OPB_MIN) cur_min[u] = STAB ? lds64(smem_saddr + p.soff_tab_min[g] + 8 * rel[u]) : G.acc_min[rel[u]];
                                    if (ops & OPB_MAX) cur_max[u] = STAB ? lds64(smem_saddr + p.soff_tab_max[g] + 8 * rel[u]) : G.acc_max[rel[u]];
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; u++) {
                            if (act[u]) {
                                if (ops & OPB_SUM) {
                                    uint64_t* a = STAB ? (uint64_t*)(smem + p.soff_tab_sum[g]) + rel[u] : G.acc_sum + rel[u];
                                    if (G.kind == TAGG_F64) atomicAdd((double*)a, c2f(code[u]));
                                    else atomicAdd((unsigned long long*)a, (unsigned long long)code_to_bits(G.kind, code[u]));
                                }
                                if (nibf) {
                                    const uint32_t fx = nfb[u] >> 4, fn = nfb[u] & 15u, qx = nq[u], qn = 15u - nq[u];
                                    if ((ops & OPB_MAX) && qx >= fx) atomicMax((unsigned long long*)(G.acc_max + rel[u]), (unsigned long long)code[u]);
                                    if ((ops & OPB_MIN) && qn >= fn) atomicMax((unsigned long long*)(G.acc_min + rel[u]), (unsigned long long)~code[u]);
                                    // (both nibbles are raised whatever the ops: the byte is also the bucket's existence record)
                                    const uint32_t nb8 = (max(fx, qx) << 4) | max(fn, qn);
                                    if (nb8 != nfb[u]) asm volatile("st.shared.u8 [%0], %1;" ::"r"(smem_saddr + p.soff_nib + rel[u]), "r"(nb8) : "memory");
                                    continue;
                                }
                                if (filt) {  // fire-and-forget: no load sits between the filter and the global RED
                                    if (ops & OPB_MIN) {
                                        const uint32_t q = BUCKET == BK_RANK ? ~rank_x[u] : (uint32_t)((p.filt_hi[g] - code[u]) >> p.filt_shift[g]);
                                        if (q >= (uint32_t)cur_min[u]) {
                                            if (q > (uint32_t)cur_min[u]) asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(smem_saddr + p.soff_tab_min[g] + 4 * rel[u]), "r"(q) : "memory");
                                            atomicMax((unsigned long long*)(G.acc_min + rel[u]), (unsigned long long)~code[u]);
                                        }
                                    }
                                    if (ops & OPB_MAX) {
                                        const uint32_t q = BUCKET == BK_RANK ? rank_x[u] : (uint32_t)((code[u] - p.filt_lo[g]) >> p.filt_shift[g]);
                                        if (q >= (uint32_t)cur_max[u]) {
                                            if (q > (uint32_t)cur_max[u]) asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(smem_saddr + p.soff_tab_max[g] + 4 * rel[u]), "r"(q) : "memory");
                                            atomicMax((unsigned long long*)(G.acc_max + rel[u]), (unsigned long long)code[u]);
                                        }
                                    }
                                    continue;
                                }
                                if ((ops & OPB_MIN) && cur_min[u] < ~code[u]) {
                                    if (STAB) atomicMax((unsigned long long*)(smem + p.soff_tab_min[g]) + rel[u], (unsigned long long)~code[u]);
                                    else if (__ldcg(G.acc_min + rel[u]) < ~code[u]) atomicMax((unsigned long long*)(G.acc_min + rel[u]), (unsigned long long)~code[u]);
                                }
                                if ((ops & OPB_MAX) && cur_max[u] < code[u]) {
                                    if (STAB) atomicMax((unsigned long long*)(smem + p.soff_tab_max[g]) + rel[u], (unsigned long long)code[u]);
                                    else if (__ldcg(G.acc_max + rel[u]) < code[u]) atomicMax((unsigned long long*)(G.acc_max + rel[u]), (unsigned long long)code[u]);
                                }
                            }
                        }
                    }
                }
            };
            // the non-negative fast path is instantiated where it pays: CT shapes and histograms
            // (only where a value is converted to f64: sums and histogram keys; min / max never need it)
            constexpr bool HAS_POS = ((SH::BOPS >= 0 && (SH::BOPS & OPB_SUM)) || (SH::ROPS >= 0 && (SH::ROPS & OPB_SUM)) || BUCKET == BK_HIST) && BUCKET != BK_RANK;
            auto run = [&](auto U_, auto CHECK_, const uint32_t* dl, const bool* act_in) {
                if (HAS_POS && fpos) process(U_, CHECK_, std::true_type{}, dl, act_in);
                else process(U_, CHECK_, std::false_type{}, dl, act_in);
            };
            using I1 = std::integral_constant<int, 1>;
            using I2 = std::integral_constant<int, 2>;
            using I4 = std::integral_constant<int, 4>;

            if (COMPACT) {
                // ---- phase 2: compact the set bits of the 8 words into the warp's queue ----------------
                // Lane l owns documents [8l, 8l + 8) of the warp's 256 — byte l & 3 of mask word l >> 2: one POPC per lane and a
                // five-step scan place all of them, in document order.  (A ballot-style rank per mask word costs two POPCs per
                // word on the quarter-rate pipe: 17 % of C2's instructions, profiles/r2_ncu_hot_C2.txt.)
                static_assert(ST_WORDS_PER_WARP == 8, "the queue compaction assumes 8 documents per lane");
                const uint32_t mw = __shfl_sync(0xffffffffu, m, lane >> 2);
                const uint32_t mb = (mw >> ((lane & 3u) * 8u)) & 0xffu;
                const uint32_t mc = __popc(mb);
                uint32_t incl = mc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (uint32_t)o) incl += t;
                }
                const uint32_t nq = __shfl_sync(0xffffffffu, incl, 31);
                {
                    uint32_t at = q_saddr + 2 * (incl - mc);
                    const uint32_t d0 = warp * ST_DOCS_PER_WARP + lane * 8;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        if (mb & (1u << k)) {
                            asm volatile("st.shared.u16 [%0], %1;" ::"r"(at), "h"((uint16_t)(d0 + k)) : "memory");
                            at += 2;
                        }
                    }
                }
                matched += nq;
                __syncwarp();
                // ---- phase 3: full warps drain the queue: branch-free batches of 64, then the ragged tail --
                uint32_t j0 = 0;
                for (; j0 + 64 <= nq; j0 += 64) {
                    uint32_t dl[2];
                    uint16_t d0, d1;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(d0) : "r"(q_saddr + 2 * (j0 + lane)));
                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(d1) : "r"(q_saddr + 2 * (j0 + 32 + lane)));
                    dl[0] = d0; dl[1] = d1;
                    run(I2{}, std::false_type{}, dl, nullptr);
                }
                if (nq - j0 > 32) {  // 33..63 left: one two-deep batch with a ragged second half
                    uint32_t dl[2];
                    bool act[2];
                    act[0] = true;
                    act[1] = j0 + 32 + lane < nq;
                    uint16_t d0, d1 = 0;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(d0) : "r"(q_saddr + 2 * (j0 + lane)));
                    if (act[1]) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(d1) : "r"(q_saddr + 2 * (j0 + 32 + lane)));
                    dl[0] = d0; dl[1] = d1;
                    run(I2{}, std::true_type{}, dl, act);
                } else if (j0 < nq) {
                    uint16_t d0 = 0;
                    const bool a0 = j0 + lane < nq;
                    if (a0) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(d0) : "r"(q_saddr + 2 * (j0 + lane)));
                    if (BUCKET == BK_RANK) {  // one ragged instantiation only (instruction-cache footprint)
                        uint32_t dl[2] = {d0, 0};
                        bool act[2] = {a0, false};
                        run(I2{}, std::true_type{}, dl, act);
                    } else {
                        uint32_t dl[1] = {d0};
                        bool act[1] = {a0};
                        run(I1{}, std::true_type{}, dl, act);
                    }
                }
            } else {
                // nothing narrows the doc stream: every document of the tile is matched (ragged only in a
                // segment's last tile)
                const uint32_t wbase = warp * ST_DOCS_PER_WARP;
                if (CTROOT && n_valid == ST_TILE) {
                    // The lane's 8 values sit 32 values = nb words apart: one address, one shift amount and one mask pair
                    // serve all of them (the generic unpack recomputes them per value).  Width and sign class are uniform
                    // over the tile, so the loop is instantiated for each instead of predicated.
                    matched += ST_DOCS_PER_WARP;
                    rseen = true;
                    fseen = true;
                    const TCol& c = rc[0];
                    const uint32_t bit0 = (wbase + lane) * c.nb, sh = bit0 & 31u, stepb = c.nb * 4;
                    const uint32_t a0 = c.saddr + ((bit0 >> 5) << 2);
                    auto dense = [&](auto WIDE_, auto POS_) {
                        constexpr bool WIDE = decltype(WIDE_)::value, POS = decltype(POS_)::value;
                        uint32_t a = a0;
#pragma unroll
                        for (int k = 0; k < ST_WORDS_PER_WARP; k++, a += stepb) {
                            const uint32_t w0 = lds32(a), w1 = lds32(a + 4);
                            const uint32_t lo = __funnelshift_r(w0, w1, sh) & c.mlo;
                            uint32_t hi = 0;
                            if (WIDE) hi = __funnelshift_r(w1, lds32(a + 8), sh) & c.mhi;
                            const uint64_t d = ((uint64_t)hi << 32) | lo;
                            if (SH::ROPS & OPB_SUM) {
                                const double v = POS ? __longlong_as_double((long long)(d + rbase)) : code_to_f64(d + cminv);
                                rsum[0] = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)rsum[0]), v));
                            }
                            if (SH::ROPS & OPB_MIN) dmin = d < dmin ? d : dmin;
                            if (SH::ROPS & OPB_MAX) dmax = d > dmax ? d : dmax;
                        }
                    };
                    if (c.nb > 32) { if (fpos) dense(std::true_type{}, std::true_type{}); else dense(std::true_type{}, std::false_type{}); }
                    else { if (fpos) dense(std::false_type{}, std::true_type{}); else dense(std::false_type{}, std::false_type{}); }
                } else if (n_valid == ST_TILE) {
                    matched += ST_DOCS_PER_WARP;
#pragma unroll
                    for (int j0 = 0; j0 < ST_WORDS_PER_WARP; j0 += 4) {
                        uint32_t dl[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) dl[u] = wbase + (j0 + u) * 32 + lane;
                        run(I4{}, std::false_type{}, dl, nullptr);
                    }
                } else {
#pragma unroll 1
                    for (int j = 0; j < ST_WORDS_PER_WARP; j++) {
                        uint32_t mj = __shfl_sync(0xffffffffu, m, j);
                        matched += __popc(mj);
                        uint32_t dl[1];
                        bool act[1];
                        dl[0] = wbase + j * 32 + lane;
                        act[0] = (mj >> lane) & 1u;
                        run(I1{}, std::true_type{}, dl, act);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_s(empty_saddr + 8 * stage);  // this warp is done with the stage
            if (++stage == S) { stage = 0; parity ^= 1u; }
        }

        if (CTROOT) fold_root();
        if (BUCKET == BK_TERMS && bad_key && p.overflow_flag) *p.overflow_flag = 4u;
        if (BUCKET == BK_RANK && wtail) flush_tail();  // the last partial buffer
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
                const uint32_t ops = g == 0 ? ops_r0 : G.ops;
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
                    if (ops & OPB_SUM) {
                        if (G.kind == TAGG_F64) atomicAdd((double*)G.acc_sum, __longlong_as_double((long long)s));
                        else atomicAdd((unsigned long long*)G.acc_sum, (unsigned long long)s);
                        *G.seen_sum = 1;
                    }
                    if (ops & OPB_MIN) { atomicMax((unsigned long long*)G.acc_min, (unsigned long long)mn); *G.seen_min = 1; }
                    if (ops & OPB_MAX) { atomicMax((unsigned long long*)G.acc_max, (unsigned long long)mx); *G.seen_max = 1; }
                }
            }
        }
    }

    if (!STAB && BUCKET == BK_TERMS && p.present_from_nib) {  // a touched bucket's level byte is never zero
        __syncthreads();
        const uint8_t* nb = smem + p.soff_nib;
        for (uint32_t i = tid; i < (uint32_t)p.dom_size; i += blockDim.x)
            if (nb[i]) p.present_out[i] = 1;
    }
    if (!STAB && BUCKET == BK_TERMS && p.soff_present_bits) {
        __syncthreads();
        const uint32_t* bm = (const uint32_t*)(smem + p.soff_present_bits);
        for (uint32_t i = tid; i < (uint32_t)p.dom_size; i += blockDim.x)
            if ((bm[i >> 5] >> (i & 31)) & 1u) p.present_out[i] = 1;
    }
    if (BUCKET == BK_RANK && p.side_dom) {
        __syncthreads();
        const uint32_t* sc = (const uint32_t*)(smem + p.soff_side_count);
        for (uint32_t i = tid; i < p.side_dom; i += blockDim.x) {
            const uint32_t v = sc[i];
            if (v) { atomicAdd((unsigned long long*)(p.side_count_acc + i), (unsigned long long)v); p.side_present[i] = 1; }
        }
    }
    if (STAB) {
        // merge the CTA's private tables into the global ones; count table 0 always exists in STAB mode
        // (hidden when the plan has no count) and is the record of which buckets exist
        __syncthreads();
        const uint32_t* sc0 = (const uint32_t*)(smem + p.soff_tab_count[0]);
        for (uint64_t i = tid; i < p.dom_size; i += blockDim.x) {
            uint32_t v = sc0[i];
            if (v) {
                if (p.n_bcounts > 0) atomicAdd((unsigned long long*)(p.bcount_acc[0] + i), (unsigned long long)v);
                if (!p.present_out[i]) p.present_out[i] = 1;
            }
        }
        if (p.n_bcounts > 1) {
            const uint32_t* sc = (const uint32_t*)(smem + p.soff_tab_count[1]);
            for (uint64_t i = tid; i < p.dom_size; i += blockDim.x) {
                uint32_t v = sc[i];
                if (v) atomicAdd((unsigned long long*)(p.bcount_acc[1] + i), (unsigned long long)v);
            }
        }
#pragma unroll
        for (int g = 0; g < NBG; g++) {
            const SGroup& G = p.bgroups[g];
            const uint32_t ops = (g == 0 && SH::BOPS >= 0) ? (uint32_t)SH::BOPS : G.ops;
            if (ops & OPB_SUM) {
                const uint64_t* ss = (const uint64_t*)(smem + p.soff_tab_sum[g]);
                for (uint64_t i = tid; i < p.dom_size; i += blockDim.x) {
                    uint64_t v = ss[i];
                    if (G.kind == TAGG_F64) { if (v != F64_NEG_ZERO_BITS) atomicAdd((double*)(G.acc_sum + i), __longlong_as_double((long long)v)); }
                    else if (v) atomicAdd((unsigned long long*)(G.acc_sum + i), (unsigned long long)v);
                }
            }
            if (p.tab_filt) continue;  // filter tables: the exact extremes went to the global table directly
            if (ops & OPB_MIN) {
                const uint64_t* ss = (const uint64_t*)(smem + p.soff_tab_min[g]);
                for (uint64_t i = tid; i < p.dom_size; i += blockDim.x) {
                    uint64_t v = ss[i];
                    if (v && __ldcg(G.acc_min + i) < v) atomicMax((unsigned long long*)(G.acc_min + i), (unsigned long long)v);
                }
            }
            if (ops & OPB_MAX) {
                const uint64_t* ss = (const uint64_t*)(smem + p.soff_tab_max[g]);
                for (uint64_t i = tid; i < p.dom_size; i += blockDim.x) {
                    uint64_t v = ss[i];
                    if (v && __ldcg(G.acc_max + i) < v) atomicMax((unsigned long long*)(G.acc_max + i), (unsigned long long)v);
                }
            }
        }
    }
}


typedef void (*stream_fn)(const SParams);
// shape families, one translation unit each (stream_inst_*.cu)
stream_fn stream_pick_rt_none_terms(int bucket, int nbg, int nrg, bool compact, bool stab);
stream_fn stream_pick_rt_hist(int nbg, int nrg, bool compact, bool stab);
stream_fn stream_pick_ct_terms(uint32_t bops0, bool compact, bool stab, bool filt);
stream_fn stream_pick_ct_hist_rank(int bucket, bool rank_linear, bool compact, bool stab);
stream_fn stream_pick_ct_root(uint32_t rops0, bool compact);

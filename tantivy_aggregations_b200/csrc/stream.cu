// stream.cu — the streaming fast path: K1 (scan_reduce), K2 (terms), K3 (histogram), K4 (percentile rank bins).
//
// One persistent launch covers every segment of the call.  The unit of work is a TILE of 2048
// consecutive documents of one segment.  For each tile the bytes of every referenced column
// (256 * num_bits bytes, contiguous and 16-byte aligned in the bit-packed layout) and 256 bytes of
// every bitset (docset, filter_agg docsets, delete bitset — a page-locked HOST bitset is pulled in place
// over PCIe) are staged into shared memory by TMA bulk copies (cp.async.bulk + mbarrier) through a 2-4
// stage ring by a producer warp, so HBM is read exactly once, fully coalesced, while 8 consumer warps per
// group unpack the previous tile from shared memory.
//
// Per tile each consumer warp owns 256 documents = 8 bitset words:
//   phase 1  lanes 0..7 AND the staged bitset words (docset, ~deletes, filter docsets) of "their"
//            word into a match mask; value predicates (post_filter / COLUMN_RANGE) are folded in — narrow
//            columns on packed 64-bit windows in the delta domain, wide ones per document with ballots;
//   phase 2  the set bits are COMPACTED into a per-warp queue of document indices;
//   phase 3  full warps drain the queue: only matched documents pay the unpack of key / value
//            columns and the table updates (the reference pays a hash probe per matched doc,
//            terms.rs:127-132; the doc-stream narrowing is filter.rs:100-122 / post_filter.rs:245-249).
// With nothing narrowing the doc stream (AllQuery, no deletes) phase 2 is skipped and the unpack is
// strength-reduced (a lane's 8 values sit exactly num_bits words apart).
//
// Root metrics live in registers (CT shapes: on the packed deltas) and are reduced by warp shuffles.  Bucket
// metrics go, depending on what fits the 227 KB of shared memory, to CTA-private exact tables, to CTA-private
// u32 filter tables in front of exact global cells, or to global tables (L2 RED atomics) behind a 4-bit level
// filter and a presence bitmap — see SParams and DESIGN.md §4.  The kernel is a template over the bucket mode
// and the number of root / bucket column groups so every column descriptor lives in registers; the hot
// configurations get compile-time op masks (CT shapes).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "stream_kernel.cuh"

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
// compile-time-op-mask (CT) shapes for the hot configurations, runtime-op-mask (RT) shapes for any flat plan
static stream_fn pick_kernel(int bucket, int nbg, int nrg, bool compact, bool stab, uint32_t bops0, uint32_t rops0, bool r0_f64, bool rank_linear, bool filt, int n_bcounts) {
    stream_fn fn = nullptr;
    if (bucket == BK_TERMS && nbg == 1 && nrg == 0 && n_bcounts <= 1) fn = stream_pick_ct_terms(bops0, compact, stab, filt);
    if (!fn && bucket == BK_HIST && nbg == 0 && nrg == 0 && n_bcounts <= 1) fn = stream_pick_ct_hist_rank(BK_HIST, false, compact, stab);  // histogram_agg_f64(field, interval, count_agg())
    if (!fn && bucket == BK_RANK) fn = stream_pick_ct_hist_rank(BK_RANK, rank_linear, compact, stab);
    if (!fn && bucket == BK_NONE && nrg == 1 && r0_f64) fn = stream_pick_ct_root(rops0, compact);
    if (fn) return fn;
    if (bucket == BK_HIST) return stream_pick_rt_hist(nbg, nrg, compact, stab);
    return stream_pick_rt_none_terms(bucket, nbg, nrg, compact, stab);
}

static bool j_monotone(const std::vector<uint64_t>& b) {
    for (size_t i = 1; i < b.size(); i++)
        if (b[i] < b[i - 1]) return false;
    return true;
}

// hist_bounds[j] = smallest code c <= code(+inf) that is not skipped and whose ordinal is >= dom_min + j (code(+inf) + 1
// if none), j = 0..dom_size.  The exact expression floor((k - start) / interval) (histogram.rs:146) is monotone in the
// order-preserving code, so each boundary is found by a local search around code(start + o * interval).
static bool hist_boundaries(double start, double interval, uint64_t dom_min, uint64_t dom_size, std::vector<uint64_t>& B) {
    const uint64_t top = f64_to_code_h(INFINITY);
    auto reaches = [&](uint64_t c, uint64_t o) { uint64_t ord; return c <= top && hist_ord_h(c, start, interval, &ord) && ord >= o; };
    B.assign(dom_size + 1, 0);
    for (uint64_t j = 0; j <= dom_size; j++) {
        const uint64_t o = dom_min + j;
        if (!reaches(top, o)) { B[j] = top + 1; continue; }
        uint64_t g = f64_to_code_h(start + (double)o * interval);
        if (g > top) g = top;
        uint64_t lo, hi;  // !reaches(lo), reaches(hi)
        if (reaches(g, o)) {
            hi = g;
            lo = g;
            uint64_t step = 1;
            bool bottom = false;
            while (true) {
                if (lo < step) { bottom = true; break; }
                lo -= step;
                if (!reaches(lo, o)) break;
                hi = lo;
                step <<= 1;
            }
            if (bottom) {
                if (reaches(0, o)) { B[j] = 0; continue; }
                lo = 0;
            }
        } else {
            lo = g;
            hi = g;
            uint64_t step = 1;
            while (true) {
                hi = top - hi < step ? top : hi + step;
                if (reaches(hi, o)) break;
                lo = hi;
                step <<= 1;
            }
        }
        while (hi - lo > 1) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (reaches(mid, o)) hi = mid; else lo = mid;
        }
        B[j] = hi;
    }
    return j_monotone(B);
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

// One streaming launch over the members of the plan's top tuple that are not handled yet: the root
// metrics (first launch only) and the first streamable bucket node.  Returns 1 if it launched (the
// members it covered are flagged in es.skip), 0 if nothing is left that streams, < 0 on error.
static int stream_launch(ExecState& es, bool first_launch) {
    const PlanMeta& m = *es.meta;
    if (es.segs.empty()) return 0;
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
        if (type != PR_FILTER) sp.n_vpreds++;
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
    // percentiles first: a histogram over the same column can ride along in that pass
    std::stable_sort(members.begin(), members.end(), [&](uint32_t a, uint32_t b) {
        return (m.nodes[a].op == TAGG_OP_PERCENTILES) > (m.nodes[b].op == TAGG_OP_PERCENTILES);
    });
    std::vector<uint32_t> covered;
    for (uint32_t mem : members) {
        if (es.skip[mem]) continue;  // handled by an earlier launch
        if (std::find(covered.begin(), covered.end(), mem) != covered.end()) continue;  // fused into another member's pass
        const tagg_node& nd = m.nodes[mem];
        if (nd.op == TAGG_OP_COUNT) {
            if (!first_launch || sp.n_root_counts >= 2) continue;
            sp.root_count_acc[sp.n_root_counts++] = (uint64_t*)(es.arena + es.slots[m.slot_of[mem]].off_acc);
            covered.push_back(mem);
        } else if (nd.op == TAGG_OP_SUM || nd.op == TAGG_OP_MIN || nd.op == TAGG_OP_MAX) {
            if (!first_launch) continue;
            Shape save = sh;
            int save_n = n_rgroups;
            if (add_fold(es, sh, sp.rgroups, n_rgroups, ST_MAXRG, (int)mem)) covered.push_back(mem);
            else { sh = save; n_rgroups = save_n; }
        } else if (nd.op == TAGG_OP_PERCENTILES) {
            // K4 on the streaming path: rank bins between sampled thresholds (pct.cu)
            if (bucket_mode != BK_NONE || nd.multi || n_bgroups > 0) continue;
            const int k = m.pct_of[mem];
            Shape save = sh;
            const int scol = sh.stage_col(m.col_slot[mem]);
            if (scol < 0) { sh = save; continue; }
            const int rc = pct_rank_plan(es, mem, k);
            if (rc < 0) return rc;
            if (rc == 0) { sh = save; continue; }
            const ExecState::RankState& R = es.rank[k];
            bucket_mode = BK_RANK;
            sp.key_scol = scol;
            sp.dom_min = 0;
            sp.dom_size = R.n_bins;
            sp.rank_lo = R.lo; sp.rank_span = R.span; sp.rank_shift = R.shift; sp.rank_mul = R.mul;
            sp.rank_linear = R.linear ? 1u : 0u; sp.rank_flo = R.f_lo; sp.rank_fscale = R.f_scale;
            sp.tail_codes = R.d_tail; sp.tail_count = R.d_tail_count; sp.tail_cap = R.tail_cap;
            sp.overflow_flag = (uint32_t*)(es.arena + es.off_overflow);
            sp.present = R.d_present;
            sp.n_bcounts = 1;
            sp.bcount_acc[0] = R.d_count;
            SGroup& G = sp.bgroups[n_bgroups++];
            memset(&G, 0, sizeof(G));
            G.scol = scol;
            G.kind = TAGG_F64;
            G.ops = OPB_MIN | OPB_MAX;
            G.acc_min = R.d_min;
            G.acc_max = R.d_max;
            covered.push_back(mem);
            // histogram_agg_f64(same column, interval, count_agg()) in the same tuple: fused (BASELINE config C3)
            for (uint32_t h : members) {
                const tagg_node& hn = m.nodes[h];
                if (es.skip[h] || hn.op != TAGG_OP_HISTOGRAM || hn.kind != TAGG_F64 || hn.multi || m.col_slot[h] != m.col_slot[mem]) continue;
                const ScopeLayout& HL = es.scopes[m.own_scope[h]];
                if (HL.mode != SCOPE_DENSE || HL.dom_size > 256 || HL.capacity != HL.dom_size) continue;
                if (m.end[h] != h + 2 || m.nodes[h + 1].op != TAGG_OP_COUNT) continue;
                sp.side_dom = (uint32_t)HL.dom_size;
                sp.side_dom_min = HL.dom_min;
                sp.f0 = hn.f0;
                sp.f1 = hn.f1;
                sp.side_count_acc = (uint64_t*)(es.arena + es.slots[m.slot_of[h + 1]].off_acc);
                sp.side_present = es.arena + HL.off_present;
                covered.push_back(h);
                break;
            }
        } else if (nd.op == TAGG_OP_TERMS || nd.op == TAGG_OP_HISTOGRAM) {
            if (bucket_mode != BK_NONE || nd.multi) continue;  // one bucket node per launch; multi-valued: generic kernel
            if (nd.op == TAGG_OP_HISTOGRAM && nd.kind != TAGG_F64) continue;  // date_histogram (integer keys): generic kernel
            const int sc = m.own_scope[mem];
            const ScopeLayout& L = es.scopes[sc];
            if (L.mode != SCOPE_DENSE) continue;
            // the sub-tree must be count / sum / min / max leaves on single-valued columns
            uint32_t sub = mem + 1;
            std::vector<uint32_t> leaves;
            if (m.nodes[sub].op == TAGG_OP_TUPLE) {
                for (uint32_t c = sub + 1; c < m.end[sub]; c = m.end[c]) leaves.push_back(c);
            } else {
                leaves.push_back(sub);
            }
            Shape save = sh;
            int save_nb = n_bgroups;
            bool ok = sh.stage_col(m.col_slot[mem]) >= 0;
            int n_bc = 0;
            uint64_t* bc_acc[2] = {nullptr, nullptr};
            for (uint32_t lf : leaves) {
                if (!ok) break;
                const tagg_node& ln = m.nodes[lf];
                if (ln.op == TAGG_OP_COUNT) {
                    if (n_bc >= 2) ok = false;
                    else bc_acc[n_bc++] = (uint64_t*)(es.arena + es.slots[m.slot_of[lf]].off_acc);
                } else if (ln.op == TAGG_OP_SUM || ln.op == TAGG_OP_MIN || ln.op == TAGG_OP_MAX) {
                    ok = add_fold(es, sh, sp.bgroups, n_bgroups, ST_MAXBG, (int)lf);
                } else {
                    ok = false;
                }
            }
            if (!ok) { sh = save; n_bgroups = save_nb; continue; }
            bucket_scope = sc;
            bucket_mode = nd.op == TAGG_OP_TERMS ? BK_TERMS : BK_HIST;
            sp.key_scol = sh.stage_col(m.col_slot[mem]);
            sp.dom_min = L.dom_min;
            sp.dom_size = L.dom_size;
            sp.f0 = nd.f0;
            sp.f1 = nd.f1;
            sp.present = es.arena + L.off_present;
            sp.n_bcounts = n_bc;
            sp.bcount_acc[0] = bc_acc[0];
            sp.bcount_acc[1] = bc_acc[1];
            covered.push_back(mem);
        }
    }
    if (covered.empty()) return 0;

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
    sp.soff_bits = off;  // (the bitset slots are sized below, once the segments' flags are known)

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
        {
            bool fpos = true;
            for (int c = 0; c < sp.n_cols; c++) {
                const DevColumn& col = hs.cols[sh.staged[c]];
                if (col.kind == TAGG_F64 && col.n_values && (col.min_value < 0x8000000000000000ull || col.max_value > 0xFFF0000000000000ull)) fpos = false;
            }
            if (fpos) d.flags |= SF_FPOS;
        }
        const uint32_t padded = (uint32_t)((((uint64_t)hs.max_doc + ST_TILE - 1) / ST_TILE) * (ST_TILE / 8));
        if (hs.main.kind == DS_BITSET) {
            d.flags |= SF_MAIN_BITS; d.bits_ptr[0] = (const uint8_t*)hs.main.words; narrowing = true;
            d.bits_len[0] = hs.main.n ? (uint32_t)hs.main.n : padded;
        }
        if (hs.has_deletes) { d.flags |= SF_DELETES; d.bits_ptr[1] = (const uint8_t*)hs.deleted; d.bits_len[1] = padded; narrowing = true; }
        for (int pi = 0; pi < sp.n_preds; pi++) {
            int src = pred_src[pi];
            if (src == -2) { d.pred_lo[pi] = hs.main.lo; d.pred_hi[pi] = hs.main.hi; }
            else if (src == -1) { d.pred_lo[pi] = pred_const[pi].first; d.pred_hi[pi] = pred_const[pi].second; }
            else {
                const DevDocset& fd = hs.filters[src];
                if (sp.pred_type[pi] == PR_RANGE) { d.pred_lo[pi] = fd.lo; d.pred_hi[pi] = fd.hi; }
                else if (fd.kind == DS_BITSET) { d.flags |= SF_PRED_BITS0 << pi; d.bits_ptr[2 + pi] = (const uint8_t*)fd.words; d.bits_len[2 + pi] = fd.n ? (uint32_t)fd.n : padded; }
                else if (fd.kind != DS_ALL) d.flags |= SF_PRED_NONE0 << pi;
            }
        }
    }
    {   // bitset slots of a stage: slot b sits at 256 * b; only as many as the highest slot any segment uses (shared
        // memory decides how many consumer groups fit — C5 stages no bitset and gains its third group from this)
        uint32_t used = 0;
        for (auto& d : descs) used |= d.flags & 0xffu;
        const uint32_t n_slots = used ? 32u - (uint32_t)__builtin_clz(used) : 0u;
        sp.stage_bytes = (sp.soff_bits + 256 * n_slots + 127) & ~127u;
        if (sp.stage_bytes == 0) sp.stage_bytes = 128;
    }
    for (auto& d : descs) {
        d.tile_bytes = (ST_TILE / 8) * __builtin_popcount(d.flags & 0xffu);
        for (int c = 0; c < sp.n_cols; c++) d.tile_bytes += (ST_TILE / 8) * d.nb[c];
    }
    if (tiles_total > 0xffffffffull) return 0;
    sp.n_segs = (uint32_t)nseg;
    sp.n_tiles = (uint32_t)tiles_total;
    sp.overflow_flag = (uint32_t*)(es.arena + es.off_overflow);

    // Option flags of the bucket slots coincide with bucket existence in the flat shape: alias them
    if (bucket_mode != BK_NONE) {
        for (size_t k = 0; k < es.slots.size(); k++)
            if (m.scope_of[m.slot_node[k]] == bucket_scope) es.slots[k].off_seen = es.scopes[bucket_scope].off_present;
    }
    for (uint32_t mem : covered) es.skip[mem] = 1;
    if (sp.n_tiles == 0) return 1;
    SegDesc* d_descs = nullptr;
    d_descs = (SegDesc*)es.cache_alloc(nseg * sizeof(SegDesc));
    if (!d_descs) return -tagg_fail(TAGG_ERR_OOM, "segment table allocation failed");
    if (cudaMemcpyAsync(d_descs, es.pin(descs.data(), nseg * sizeof(SegDesc)), nseg * sizeof(SegDesc), cudaMemcpyHostToDevice, es.st) != cudaSuccess)
        return -tagg_fail(TAGG_ERR_CUDA, "segment table upload failed");
    sp.segs = d_descs;

    // shared-memory plan.  STAB (CTA-private count / sum tables) whenever the tables leave room for the
    // staging ring; then one CTA per SM with as many groups and stages as fit.
    const size_t SMEM_MAX = 225 * 1024;
    auto group_bytes = [&](uint32_t stages) {
        size_t b = (size_t)stages * sp.stage_bytes + ST_WARPS * ST_DOCS_PER_WARP * 2 + (size_t)stages * sizeof(TileDesc) + 2 * stages * 8;
        b = (b + 127) & ~(size_t)127;
        if (bucket_mode == BK_RANK) b += ST_WARPS * ST_TBUF * 8;  // per-warp buffers of out-of-range codes
        return b;
    };
    size_t table_bytes = 0;
    bool stab = false;
    if (bucket_mode != BK_NONE) {
        auto lay_tables = [&](bool filt) {
            size_t tb = 0;
            const size_t mm = filt ? 4 : 8;
            // count table 0 always exists in a shared-table kernel: it records which buckets exist
            for (int c = 0; c < std::max(sp.n_bcounts, 1); c++) { sp.soff_tab_count[c] = (uint32_t)tb; tb += ((sp.dom_size * 4 + 127) & ~127ull); }
            for (int g = 0; g < n_bgroups; g++) {
                if (sp.bgroups[g].ops & OPB_SUM) { sp.soff_tab_sum[g] = (uint32_t)tb; tb += ((sp.dom_size * 8 + 127) & ~127ull); }
                if (sp.bgroups[g].ops & OPB_MIN) { sp.soff_tab_min[g] = (uint32_t)tb; tb += ((sp.dom_size * mm + 127) & ~127ull); }
                if (sp.bgroups[g].ops & OPB_MAX) { sp.soff_tab_max[g] = (uint32_t)tb; tb += ((sp.dom_size * mm + 127) & ~127ull); }
            }
            return tb;
        };
        bool has_mm = false;
        for (int g = 0; g < n_bgroups; g++) has_mm = has_mm || (sp.bgroups[g].ops & (OPB_MIN | OPB_MAX));
        size_t tb = lay_tables(false);
        // exact tables leave no room for a third consumer group but filter tables do: take the filter tables
        // (rank bins: always — a new extreme per bin is frequent enough that the 64-bit shared atomicMax, a CAS loop,
        // shows up; the filter path is a 32-bit RED.MAX plus a global RED)
        if (has_mm && sp.dom_size <= (1u << 20) && (bucket_mode == BK_RANK || tb + 3 * group_bytes(2) > SMEM_MAX)) {
            const size_t tf = lay_tables(true);
            if (tf + 3 * group_bytes(2) <= SMEM_MAX || bucket_mode == BK_RANK) { sp.tab_filt = 1; tb = tf; }
            else tb = lay_tables(false);
        }
        // (the CTA-private count tables are u32: the persistent grid spreads the call's tiles evenly over >= 148 CTAs, so a
        //  CTA counts at most ~1/148 of the call's documents — guard with a factor of two to spare)
        uint64_t call_docs = 0;
        for (auto& hs : es.hsegs) call_docs += hs.max_doc;
        const bool counts_fit_u32 = call_docs < (1ull << 32) * 64;
        if (tb > 0 && sp.dom_size <= (1u << 20) && tb + group_bytes(2) <= SMEM_MAX && counts_fit_u32) { stab = true; table_bytes = tb; }
        else sp.tab_filt = 0;
        if (sp.tab_filt) {
            for (int g = 0; g < n_bgroups; g++) {
                uint64_t glo = ~0ull, ghi = 0;
                for (auto& hs : es.hsegs) {
                    const DevColumn& col = hs.cols[sh.staged[sp.bgroups[g].scol]];
                    if (!col.n_values) continue;
                    glo = std::min(glo, col.min_value);
                    ghi = std::max(ghi, col.max_value);
                }
                if (glo > ghi) { glo = 0; ghi = 0; }
                const uint64_t span = ghi - glo;
                const uint32_t bits = span ? 64 - (uint32_t)__builtin_clzll(span) : 0;
                sp.filt_lo[g] = glo;
                sp.filt_hi[g] = ghi;
                sp.filt_shift[g] = bits > 32 ? bits - 32 : 0;
            }
        }
    }
    // histogram with few buckets (the launch's bucket node, or the one fused into a percentile pass): exact code
    // boundaries of the ordinals
    {
        const bool own = bucket_mode == BK_HIST && stab && sp.dom_size <= 256 && tiles_total >= 2048;
        const bool side = bucket_mode == BK_RANK && sp.side_dom > 0;
        if (own || side) {
            std::vector<uint64_t> B;
            const uint64_t hdom_min = own ? sp.dom_min : sp.side_dom_min, hdom = own ? sp.dom_size : sp.side_dom;
            if (hist_boundaries(sp.f0, sp.f1, hdom_min, hdom, B)) {
                uint64_t* d_b = nullptr;
                d_b = (uint64_t*)es.cache_alloc(B.size() * 8);
                if (!d_b) return -tagg_fail(TAGG_ERR_OOM, "histogram boundary table allocation failed");
                if (cudaMemcpyAsync(d_b, es.pin(B.data(), B.size() * 8), B.size() * 8, cudaMemcpyHostToDevice, es.st) != cudaSuccess)
                    return -tagg_fail(TAGG_ERR_CUDA, "histogram boundary table upload failed");
                sp.hist_bounds = d_b;
                sp.hist_inv = 1.0 / sp.f1;
                sp.soff_hist_bounds = (uint32_t)table_bytes;
                table_bytes += (B.size() * 8 + 127) & ~(size_t)127;
                if (side) { sp.soff_side_count = (uint32_t)table_bytes; table_bytes += ((size_t)sp.side_dom * 4 + 127) & ~(size_t)127; }
            } else if (side) {
                return -tagg_fail(TAGG_ERR_CUDA, "histogram boundaries are not monotone (internal error)");
            }
        }
    }
    // global tables that no count names: bucket existence through a CTA bitmap in shared memory
    if (bucket_mode == BK_TERMS && !stab && sp.n_bcounts == 0 && sp.dom_size <= (1u << 20)) {
        sp.soff_present_bits = 128;
        table_bytes = 128 + ((((size_t)sp.dom_size + 31) / 32 * 4 + 127) & ~(size_t)127);
    }
    // global tables with min / max on the first bucket group: 4-bit level filter in shared memory (see SParams::soff_nib)
    bool nib = false;
    static const bool no_nib = getenv("TAGG_NO_NIB") != nullptr;  // experiment switch
    if (!no_nib && bucket_mode == BK_TERMS && !stab && n_bgroups >= 1 && (sp.bgroups[0].ops & (OPB_MIN | OPB_MAX)) && sp.dom_size <= 160 * 1024) {
        uint64_t glo = ~0ull, ghi = 0;
        for (auto& hs : es.hsegs) {
            const DevColumn& col = hs.cols[sh.staged[sp.bgroups[0].scol]];
            if (!col.n_values) continue;
            glo = std::min(glo, col.min_value);
            ghi = std::max(ghi, col.max_value);
        }
        if (glo <= ghi) {
            const uint64_t span = ghi - glo;
            const uint32_t bits = span ? 64 - (uint32_t)__builtin_clzll(span) : 0;
            sp.nib_lo = glo;
            sp.nib_shift = bits > 4 ? bits - 4 : 0;
            sp.nib_hi32 = sp.nib_shift >= 32 ? 1u : 0u;
            // the level bytes also record which buckets exist: they replace the CTA bitmap
            sp.present_from_nib = sp.soff_present_bits ? 1u : 0u;
            sp.soff_present_bits = 0;
            table_bytes = 128;
            sp.soff_nib = (uint32_t)table_bytes;
            table_bytes += ((size_t)sp.dom_size + 127) & ~(size_t)127;
            nib = true;
        }
    }
    uint32_t n_groups = 1, n_stages = 3;
    if (nib) {  // tables are per CTA: one CTA per SM with as many consumer groups as fit
        const uint32_t cand[][2] = {{3, 3}, {3, 2}, {2, 3}, {2, 2}, {1, 3}, {1, 2}};
        bool ok = false;
        for (auto& c : cand)
            if (table_bytes + c[0] * group_bytes(c[1]) <= SMEM_MAX) { n_groups = c[0]; n_stages = c[1]; ok = true; break; }
        if (const char* ov = getenv("TAGG_STREAM_GS")) {  // experiment: "groups,stages"
            uint32_t g = 0, st = 0;
            if (sscanf(ov, "%u,%u", &g, &st) == 2 && g >= 1 && g <= ST_MAXGROUPS && st >= 2 && st <= ST_MAXSTAGES && table_bytes + g * group_bytes(st) <= SMEM_MAX) { n_groups = g; n_stages = st; }
        }
        if (!ok) {  // no room for the level bytes: back to the bitmap alone
            nib = false;
            sp.soff_nib = 0;
            if (sp.present_from_nib) sp.soff_present_bits = 128;
            sp.present_from_nib = 0;
            table_bytes = sp.soff_present_bits ? 128 + ((((size_t)sp.dom_size + 31) / 32 * 4 + 127) & ~(size_t)127) : 0;
        }
    }
    if (stab) {
        // more consumer warps beat a deeper ring (measured on C3: 3 groups x 2 stages 2.08 ms, 2 x 3 2.58 ms)
        const uint32_t cand[][2] = {{3, 4}, {3, 3}, {3, 2}, {2, 4}, {2, 3}, {2, 2}, {1, 4}, {1, 3}, {1, 2}};
        bool ok = false;
        for (auto& c : cand)
            if (table_bytes + c[0] * group_bytes(c[1]) <= SMEM_MAX) { n_groups = c[0]; n_stages = c[1]; ok = true; break; }
        if (const char* ov = getenv("TAGG_STREAM_GS")) {  // experiment: "groups,stages"
            uint32_t g = 0, st = 0;
            if (sscanf(ov, "%u,%u", &g, &st) == 2 && g >= 1 && g <= ST_MAXGROUPS && st >= 2 && st <= ST_MAXSTAGES && table_bytes + g * group_bytes(st) <= SMEM_MAX) { n_groups = g; n_stages = st; }
        }
        if (!ok) { stab = false; table_bytes = 0; sp.hist_bounds = nullptr; sp.tab_filt = 0; }
    }
    if (bucket_mode == BK_RANK && !stab) return -tagg_fail(TAGG_ERR_CUDA, "rank-bin tables do not fit shared memory (internal sizing error)");
    if (!stab && !nib) {
        // one group per CTA, several CTAs per SM: take the ring depth that keeps the most consumer warps resident
        n_groups = 1;
        if (table_bytes + group_bytes(2) > SMEM_MAX) return 0;
        const size_t per3 = SMEM_MAX / (table_bytes + group_bytes(3) + 1024), per2 = SMEM_MAX / (table_bytes + group_bytes(2) + 1024);
        n_stages = (per3 >= 1 && per3 >= std::min<size_t>(per2, 4)) ? 3 : 2;
        if (const char* ov = getenv("TAGG_STREAM_S")) {  // experiment: ring depth of the one-group-per-CTA shapes
            const uint32_t st = (uint32_t)atoi(ov);
            if (st >= 2 && st <= ST_MAXSTAGES && table_bytes + group_bytes(st) <= SMEM_MAX) n_stages = st;
        }
    }
    sp.n_stages = n_stages;
    sp.group_bytes = (uint32_t)group_bytes(n_stages);
    sp.table_bytes = (uint32_t)table_bytes;
    size_t smem_bytes = table_bytes + (size_t)n_groups * sp.group_bytes;
    uint8_t* present = sp.present;
    sp.present_out = present;
    if (bucket_mode != BK_NONE && (sp.n_bcounts > 0 || stab)) sp.present = nullptr;  // derived from the counts

    // compaction pays when documents are filtered out; with nothing narrowing the stream it is pure overhead
    int nrg_t = n_rgroups;
    stream_fn fn = pick_kernel(bucket_mode, n_bgroups, nrg_t, narrowing, stab, n_bgroups ? sp.bgroups[0].ops : 0u, n_rgroups ? sp.rgroups[0].ops : 0u, n_rgroups && sp.rgroups[0].kind == TAGG_F64, sp.rank_linear != 0, sp.tab_filt != 0, sp.n_bcounts);
    static std::mutex attr_mu;
    static std::vector<stream_fn> attr_done;
    {
        std::lock_guard<std::mutex> g(attr_mu);
        if (std::find(attr_done.begin(), attr_done.end(), fn) == attr_done.end()) {
            if (cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX) != cudaSuccess)
                return -tagg_fail(TAGG_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
            attr_done.push_back(fn);
        }
    }
    int threads = (int)n_groups * ST_GROUP_THREADS;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, threads, smem_bytes) != cudaSuccess || per_sm < 1)
        return -tagg_fail(TAGG_ERR_CUDA, "k_stream does not fit an SM (%zu bytes of shared memory)", smem_bytes);
    // one launch per chunk of segments: with host docsets being uploaded (on the upload stream), chunk c's kernel
    // waits only for its own uploads and runs while later chunks are still crossing PCIe
    const bool piped = !es.uploads.empty();
    cudaStream_t kst = es.st;
    for (uint32_t c = 0; c < es.n_chunks; c++) {
        const uint32_t s0 = es.chunk_begin[c], s1 = es.chunk_begin[c + 1];
        if (s0 == s1) continue;
        SParams cp = sp;
        cp.segs = d_descs + s0;
        cp.n_segs = s1 - s0;
        const uint32_t t0 = descs[s0].tile_begin;
        const uint32_t t1 = s1 < nseg ? descs[s1].tile_begin : sp.n_tiles;
        cp.n_tiles = t1 - t0;
        cp.tile_base = t0;
        if (cp.n_tiles == 0) continue;
        if (piped && cudaStreamWaitEvent(kst, es.call->chunk_ev[c], 0) != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "stream ordering failed");
        // the pass timer (tagg_result_stats kernel_ms) starts at the first kernel of the pass, behind the descriptor
        // uploads — not at the host-side preparation above
        if (es.n_launches == es.launches_at_ev0 && es.ev0) cudaEventRecord(es.ev0, kst);
        uint64_t work_units = ((uint64_t)cp.n_tiles + n_groups - 1) / n_groups;
        const uint64_t sms = (uint64_t)es.ctx->sm_count > es.reserve_sms ? (uint64_t)es.ctx->sm_count - es.reserve_sms : 1;
        uint32_t grid = (uint32_t)std::min<uint64_t>(sms * per_sm, work_units);
        fn<<<grid, threads, smem_bytes, kst>>>(cp);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "k_stream launch failed: %s", cudaGetErrorString(e));
        es.ctx->launches++;
        es.n_launches++;
    }
    if (bucket_mode != BK_NONE && sp.n_bcounts > 0 && !stab) {
        uint64_t n = es.scopes[bucket_scope].capacity;
        k_present_from_counts<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 1024), 256, 0, es.st>>>(sp.bcount_acc[0], present, n);
        es.ctx->launches++;
        es.n_launches++;
    }
    return 1;
}

// Runs as many streaming launches as the plan's top tuple needs (root metrics + one bucket node per
// launch; every launch re-reads only the columns it uses).  Members that do not stream (multi-valued
// fields, nested buckets, hashed scopes, percentiles) are left to the generic kernel: es.skip tells it
// which sub-trees are already done.  Returns 1: everything streamed, 2: partly, 0: nothing, < 0: error.
int stream_try(ExecState& es) {
    const PlanMeta& m = *es.meta;
    es.skip.assign(m.nodes.size(), 0);
    int launches = 0;
    for (int i = 0; i < 8; i++) {
        int rc = stream_launch(es, i == 0);
        if (rc < 0) return rc;
        if (rc == 0) break;
        launches++;
    }
    if (!launches) return 0;
    // which top-level members are left?
    uint32_t node = 0;
    while (node < m.nodes.size() && (m.nodes[node].op == TAGG_OP_FILTER || m.nodes[node].op == TAGG_OP_POST_FILTER)) node++;
    bool all = true;
    if (m.nodes[node].op == TAGG_OP_TUPLE) {
        for (uint32_t c = node + 1; c < m.end[node]; c = m.end[c]) all = all && es.skip[c];
    } else {
        all = es.skip[node];
    }
    return all ? 1 : 2;
}

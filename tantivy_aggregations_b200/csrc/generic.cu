// generic.cu — the tree-walking kernel: one thread per candidate document, the plan interpreted
// in pre-order exactly like the reference's monomorphised `SegmentAgg::collect` recursion
// (src/searcher.rs:41-48 -> src/tuple.rs:63-67 -> leaf collects).  It supports EVERY plan the
// API can express (nested buckets, multi-valued fields, predicates) and is the fallback when a
// plan has no streaming fast shape (stream.cu).  Still a CUDA path: there is no CPU fallback.
#include "dev.cuh"
#include "host.h"

__device__ __forceinline__ bool pred_test(const DevNode& nd, uint64_t code) {
    if (nd.pred == TAGG_PRED_RANGE) return code >= nd.u0 && code <= nd.u1;
    if (nd.pred == TAGG_PRED_LUT) {
        if (code < nd.u0) return false;
        uint64_t i = code - nd.u0;
        if (i >= nd.u1) return false;
        return (nd.lut[i >> 3] >> (i & 7)) & 1;
    }
    return true;
}

// One value folded into a SUM / MIN / MAX leaf (sum.rs:95-102, minmax.rs:97-106).
__device__ __forceinline__ void fold_value(const DevPlan* P, const DevNode& nd, uint32_t bucket, uint64_t code,
                                           uint64_t* racc, uint32_t& rseen, uint64_t pos) {
    int ri = P->slot_root_index[nd.slot];
    if (ri >= 0) {  // root scope: per-thread accumulator, reduced once at the end of the kernel
        if (nd.op == TAGG_OP_SUM) {
            if (nd.kind == TAGG_F64)
                racc[ri] = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)racc[ri]), code_to_f64(code)));
            else
                racc[ri] += code_to_bits(nd.kind, code);
        } else if (nd.op == TAGG_OP_MIN) {
            uint64_t v = ~code;
            if (v > racc[ri]) racc[ri] = v;
        } else {
            if (code > racc[ri]) racc[ri] = code;
        }
        rseen |= 1u << ri;
        return;
    }
    const DevSlot& sl = P->slots[nd.slot];
    if (nd.op == TAGG_OP_SUM) {
        if (nd.kind == TAGG_F64)
            atomicAdd((double*)(sl.acc + bucket), code_to_f64(code));
        else
            atomicAdd((unsigned long long*)(sl.acc + bucket), (unsigned long long)code_to_bits(nd.kind, code));
    } else if (sl.edge) {
        // Exact `PartialOrd` fold of minmax.rs:97-106 on f64: a NaN never replaces, but a FIRST NaN sticks; among the two
        // zeros the first one seen stays.  Record the earliest position of any value / of each zero; NaNs stay out of the
        // order-preserving min / max; k_edge_fixup settles the cell after the pass.
        const uint64_t tag = ~pos;
        if (*((volatile uint64_t*)(sl.edge + bucket)) < tag) atomicMax((unsigned long long*)(sl.edge + bucket), (unsigned long long)tag);
        if (code == CODE_NEG_ZERO || code == CODE_POS_ZERO) {
            uint64_t* z = sl.edge + (code == CODE_NEG_ZERO ? 1 : 2) * sl.edge_cap + bucket;
            if (*((volatile uint64_t*)z) < tag) atomicMax((unsigned long long*)z, (unsigned long long)tag);
        }
        if (code >= CODE_NEG_INF && code <= CODE_POS_INF) {
            uint64_t v = nd.op == TAGG_OP_MIN ? ~code : code;
            if (*((volatile uint64_t*)(sl.acc + bucket)) < v) atomicMax((unsigned long long*)(sl.acc + bucket), (unsigned long long)v);
        }
    } else {
        uint64_t v = nd.op == TAGG_OP_MIN ? ~code : code;
        // the plain read may be stale but the cell only grows: skipping when v <= stale is safe
        if (*((volatile uint64_t*)(sl.acc + bucket)) < v) atomicMax((unsigned long long*)(sl.acc + bucket), (unsigned long long)v);
    }
    if (!sl.seen[bucket]) sl.seen[bucket] = 1;
}

struct Frame {
    uint16_t end, body, node;
    uint16_t is_loop;
    uint32_t saved_bucket;
    uint64_t cur, stop;
};

__global__ void __launch_bounds__(256) k_generic(const DevPlan* __restrict__ P, const DevSegment* __restrict__ Sp,
                                                 uint64_t n_cand, uint32_t seg_index) {
    const DevSegment& S = *Sp;
    uint64_t racc[TAGG_MAX_ROOT_SLOTS];
    uint32_t rseen = 0;
    const uint64_t pos_base = (uint64_t)seg_index << EDGE_POS_BITS;
#pragma unroll
    for (int i = 0; i < TAGG_MAX_ROOT_SLOTS; i++) racc[i] = 0;
    for (uint32_t ri = 0; ri < P->n_root_slots; ri++) {  // f64 sums fold from -0.0 (x + -0.0 == x, so the first value "replaces")
        const DevNode& rn = P->nodes[P->root_slot_nodes[ri]];
        if (rn.op == TAGG_OP_SUM && rn.kind == TAGG_F64) racc[ri] = F64_NEG_ZERO_BITS;
    }

    const uint32_t n_nodes = P->n_nodes;
    const bool by_ids = S.main.kind == DS_IDS;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cand; i += stride) {
        uint32_t doc;
        if (by_ids) {
            doc = S.main.ids[i];
            // the ids come from the caller: one the segment does not have, or a list that is not strictly ascending (what a
            // tantivy scorer yields), fails the call instead of reading out of bounds / double counting
            if (doc >= S.max_doc || (i > 0 && S.main.ids[i - 1] >= doc)) {
                atomicExch(P->overflow, 5u);
                if (doc >= S.max_doc) continue;
            }
        } else {
            doc = (uint32_t)i;
            if (!docset_test(S, S.main, doc)) continue;
        }
        // searcher.rs:41-46 — deleted documents are skipped at the top loop only
        if (S.has_deletes && ((S.deleted[doc >> 5] >> (doc & 31)) & 1u)) continue;

        Frame frames[TAGG_MAX_DEPTH];
        int sp = 0;
        uint32_t bucket = 0;
        uint32_t pc = 0;
        for (;;) {
            // close every bucket frame that ends here; a multi-valued TERMS frame loops per value occurrence
            bool resumed = false;
            while (sp > 0 && frames[sp - 1].end == pc) {
                Frame& f = frames[sp - 1];
                if (f.is_loop) {
                    const DevNode& nd = P->nodes[f.node];
                    uint32_t b = INVALID_BUCKET;
                    while (++f.cur < f.stop) {
                        b = scope_lookup(P->overflow, P->scopes[nd.own_scope], f.saved_bucket, col_get(S.cols[nd.col + 1], f.cur));
                        if (b != INVALID_BUCKET) break;
                    }
                    if (b != INVALID_BUCKET) {
                        bucket = b;
                        pc = f.body;
                        resumed = true;
                        break;
                    }
                }
                bucket = f.saved_bucket;
                sp--;
            }
            if (!resumed && pc >= n_nodes) break;
            const DevNode& nd = P->nodes[pc];
            if (nd.skip) { pc = nd.end; continue; }
            switch (nd.op) {
                case TAGG_OP_TUPLE: pc++; break;
                case TAGG_OP_COUNT: {  // count.rs:53-55
                    int ri = P->slot_root_index[nd.slot];
                    if (ri >= 0) racc[ri] += 1;
                    else atomicAdd((unsigned long long*)(P->slots[nd.slot].acc + bucket), 1ull);
                    pc++;
                    break;
                }
                case TAGG_OP_SUM:
                case TAGG_OP_MIN:
                case TAGG_OP_MAX: {
                    if (!nd.multi) {
                        fold_value(P, nd, bucket, col_get(S.cols[nd.col], doc), racc, rseen, pos_base | doc);
                    } else {  // sum.rs:131-140, minmax.rs:135-145: every value of the doc
                        uint64_t a = col_get(S.cols[nd.col], doc), b = col_get(S.cols[nd.col], (uint64_t)doc + 1);
                        for (uint64_t j = a; j < b; j++) fold_value(P, nd, bucket, col_get(S.cols[nd.col + 1], j), racc, rseen, pos_base | j);
                    }
                    pc++;
                    break;
                }
                case TAGG_OP_PERCENTILES: {  // percentile.rs:87-90,119-124: every value is inserted
                    uint32_t ps = nd.aux;
                    uint64_t a, b;
                    if (!nd.multi) { a = doc; b = (uint64_t)doc + 1; }
                    else { a = col_get(S.cols[nd.col], doc); b = col_get(S.cols[nd.col], (uint64_t)doc + 1); }
                    const DevColumn& vc = S.cols[nd.multi ? nd.col + 1 : nd.col];
                    for (uint64_t j = a; j < b; j++) {
                        // warp-aggregated append: one atomic per converged group of lanes, not per value
                        const unsigned grp = __activemask();
                        const int leader = __ffs(grp) - 1;
                        unsigned long long base = 0;
                        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(P->pct_count[ps], (unsigned long long)__popc(grp));
                        base = __shfl_sync(grp, base, leader);
                        unsigned long long at = base + __popc(grp & ((1u << (threadIdx.x & 31)) - 1u));
                        if (at < P->pct_cap[ps]) {
                            P->pct_codes[ps][at] = col_get(vc, j);
                            P->pct_buckets[ps][at] = bucket;
                        } else {
                            atomicExch(P->overflow, 2u);
                        }
                    }
                    pc++;
                    break;
                }
                case TAGG_OP_TERMS: {
                    if (!nd.multi) {  // terms.rs:127-132
                        uint32_t b = scope_lookup(P->overflow, P->scopes[nd.own_scope], bucket, col_get(S.cols[nd.col], doc));
                        if (b == INVALID_BUCKET || sp >= TAGG_MAX_DEPTH) { pc = nd.end; break; }
                        Frame& f = frames[sp++];
                        f.end = nd.end; f.body = pc + 1; f.node = pc; f.is_loop = 0; f.saved_bucket = bucket;
                        bucket = b;
                        pc++;
                    } else {  // terms.rs:172-179: once per value occurrence
                        uint64_t a = col_get(S.cols[nd.col], doc), e = col_get(S.cols[nd.col], (uint64_t)doc + 1);
                        uint32_t b = INVALID_BUCKET;
                        uint64_t cur = a;
                        for (; cur < e; cur++) {
                            b = scope_lookup(P->overflow, P->scopes[nd.own_scope], bucket, col_get(S.cols[nd.col + 1], cur));
                            if (b != INVALID_BUCKET) break;
                        }
                        if (b == INVALID_BUCKET || sp >= TAGG_MAX_DEPTH) { pc = nd.end; break; }
                        Frame& f = frames[sp++];
                        f.end = nd.end; f.body = pc + 1; f.node = pc; f.is_loop = 1; f.saved_bucket = bucket;
                        f.cur = cur; f.stop = e;
                        bucket = b;
                        pc++;
                    }
                    break;
                }
                case TAGG_OP_HISTOGRAM: {  // histogram.rs:136-152
                    uint64_t ord;
                    if (!hist_ord(col_get(S.cols[nd.col], doc), nd.f0, nd.f1, &ord, nd.kind)) { pc = nd.end; break; }
                    uint32_t b = scope_lookup(P->overflow, P->scopes[nd.own_scope], bucket, ord);
                    if (b == INVALID_BUCKET || sp >= TAGG_MAX_DEPTH) { pc = nd.end; break; }
                    Frame& f = frames[sp++];
                    f.end = nd.end; f.body = pc + 1; f.node = pc; f.is_loop = 0; f.saved_bucket = bucket;
                    bucket = b;
                    pc++;
                    break;
                }
                case TAGG_OP_FILTER:  // filter.rs:100-122 == membership in the second query's docset
                    pc = docset_test(S, S.filters[nd.aux], doc) ? pc + 1 : nd.end;
                    break;
                case TAGG_OP_POST_FILTER: {
                    bool pass = false;
                    if (!nd.multi) {  // post_filter.rs:245-249
                        pass = pred_test(nd, col_get(S.cols[nd.col], doc));
                    } else {  // post_filter.rs:289-297: any value passes; the doc is collected once
                        uint64_t a = col_get(S.cols[nd.col], doc), e = col_get(S.cols[nd.col], (uint64_t)doc + 1);
                        for (uint64_t j = a; j < e && !pass; j++) pass = pred_test(nd, col_get(S.cols[nd.col + 1], j));
                    }
                    pc = pass ? pc + 1 : nd.end;
                    break;
                }
                default: pc = n_nodes; break;
            }
        }
    }

    // fold the per-thread root accumulators: warp shuffle, then one atomic per warp and slot
    const uint32_t nroot = P->n_root_slots;
    for (uint32_t ri = 0; ri < nroot; ri++) {
        const DevNode& nd = P->nodes[P->root_slot_nodes[ri]];
        uint64_t v = racc[ri];
        bool f64sum = nd.op == TAGG_OP_SUM && nd.kind == TAGG_F64;
        bool is_add = nd.op == TAGG_OP_COUNT || nd.op == TAGG_OP_SUM;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            uint64_t w = __shfl_xor_sync(0xffffffffu, v, o);
            if (f64sum) v = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)v), __longlong_as_double((long long)w)));
            else if (is_add) v += w;
            else v = v > w ? v : w;
        }
        uint32_t any = __ballot_sync(0xffffffffu, (rseen >> ri) & 1u);
        if ((threadIdx.x & 31) == 0) {
            const DevSlot& sl = P->slots[nd.slot];
            if (nd.op == TAGG_OP_COUNT) {
                if (v) atomicAdd((unsigned long long*)sl.acc, (unsigned long long)v);
            } else if (any) {
                if (f64sum) atomicAdd((double*)sl.acc, __longlong_as_double((long long)v));
                else if (is_add) atomicAdd((unsigned long long*)sl.acc, (unsigned long long)v);
                else atomicMax((unsigned long long*)sl.acc, (unsigned long long)v);
                sl.seen[0] = 1;
            }
        }
    }
}

cudaError_t launch_generic(const DevPlan* dplan, const DevSegment* dseg, uint64_t n_cand, uint32_t seg_index, int sm_count,
                           cudaStream_t stream) {
    if (n_cand == 0) return cudaSuccess;
    uint64_t blocks = (n_cand + 255) / 256;
    uint64_t cap = (uint64_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_generic<<<(unsigned)blocks, 256, 0, stream>>>(dplan, dseg, n_cand, seg_index);
    return cudaGetLastError();
}

// After the pass, per cell of an f64 MIN / MAX slot in edge mode: the reference's fold returns the FIRST collected value
// if that value is NaN (nothing is `lt` / `gt` a NaN, minmax.rs:99-102), else the extreme of the non-NaN values, and
// when that extreme is a zero and both zeros were collected, the zero that came first (-0.0 == +0.0 never replaces).
__global__ void k_edge_fixup(const DevSegment* __restrict__ segs, int col, int is_min, uint64_t* __restrict__ acc,
                             const uint8_t* __restrict__ seen, const uint64_t* __restrict__ edge, uint64_t cap) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t tag = edge[i];
        if (!tag || !seen[i]) continue;
        const uint64_t pos = ~tag;
        const uint64_t first = col_get(segs[pos >> EDGE_POS_BITS].cols[col], pos & ((1ull << EDGE_POS_BITS) - 1));
        if (first < CODE_NEG_INF || first > CODE_POS_INF) { acc[i] = is_min ? ~first : first; continue; }
        const uint64_t cur = is_min ? ~acc[i] : acc[i];
        if (cur == CODE_NEG_ZERO || cur == CODE_POS_ZERO) {
            const uint64_t nz = edge[cap + i], pz = edge[2 * cap + i];
            if (nz && pz) {
                const uint64_t pick = nz > pz ? CODE_NEG_ZERO : CODE_POS_ZERO;  // larger tag = earlier position
                acc[i] = is_min ? ~pick : pick;
            }
        }
    }
}
cudaError_t launch_edge_fixup(const DevSegment* segs, int col, int is_min, uint64_t* acc, const uint8_t* seen, const uint64_t* edge,
                              uint64_t cap, int sm_count, cudaStream_t stream) {
    uint64_t blocks = (cap + 255) / 256, lim = (uint64_t)sm_count * 8;
    if (blocks > lim) blocks = lim;
    k_edge_fixup<<<(unsigned)blocks, 256, 0, stream>>>(segs, col, is_min, acc, seen, edge, cap);
    return cudaGetLastError();
}

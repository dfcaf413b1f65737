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
                                           uint64_t* racc, uint32_t& rseen) {
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
                                                 uint64_t n_cand) {
    const DevSegment& S = *Sp;
    uint64_t racc[TAGG_MAX_ROOT_SLOTS];
    uint32_t rseen = 0;
#pragma unroll
    for (int i = 0; i < TAGG_MAX_ROOT_SLOTS; i++) racc[i] = 0;

    const uint32_t n_nodes = P->n_nodes;
    const bool by_ids = S.main.kind == DS_IDS;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cand; i += stride) {
        uint32_t doc;
        if (by_ids) {
            doc = S.main.ids[i];
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
                        fold_value(P, nd, bucket, col_get(S.cols[nd.col], doc), racc, rseen);
                    } else {  // sum.rs:131-140, minmax.rs:135-145: every value of the doc
                        uint64_t a = col_get(S.cols[nd.col], doc), b = col_get(S.cols[nd.col], (uint64_t)doc + 1);
                        for (uint64_t j = a; j < b; j++) fold_value(P, nd, bucket, col_get(S.cols[nd.col + 1], j), racc, rseen);
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
                    if (!hist_ord(col_get(S.cols[nd.col], doc), nd.f0, nd.f1, &ord)) { pc = nd.end; break; }
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

cudaError_t launch_generic(const DevPlan* dplan, const DevSegment* dseg, uint64_t n_cand, int sm_count,
                           cudaStream_t stream) {
    if (n_cand == 0) return cudaSuccess;
    uint64_t blocks = (n_cand + 255) / 256;
    uint64_t cap = (uint64_t)sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_generic<<<(unsigned)blocks, 256, 0, stream>>>(dplan, dseg, n_cand);
    return cudaGetLastError();
}

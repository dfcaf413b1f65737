// narrow.cuh — the doc-stream narrowing in front of the aggregations: main docset, delete bitset
// (searcher.rs:41-46), filter_agg docsets (filter.rs:100-122) and post_filter_agg predicates
// (post_filter.rs:245-249, 289-297), evaluated per document from the generic segment descriptors.
// Used by the kernels that are not TMA-staged (k_mterms, k_pct_sample).
#pragma once
#include "exec.h"

#define NARROW_MAXPRED 4
enum { MP_FILTER = 0, MP_RANGE = 1, MP_LUT = 2, MP_RANGE_ANY = 3, MP_LUT_ANY = 4 };
struct MPred {
    int32_t type, col, filter, pad;
    uint64_t lo, hi;
    const uint8_t* lut;
};

#ifdef __CUDACC__
__device__ __forceinline__ bool mpred_value(const MPred& pr, uint64_t code) {
    if (pr.type == MP_RANGE || pr.type == MP_RANGE_ANY) return code >= pr.lo && code <= pr.hi;
    if (code < pr.lo) return false;
    uint64_t r = code - pr.lo;
    return r < pr.hi && ((pr.lut[r >> 3] >> (r & 7)) & 1);
}


__device__ __forceinline__ bool doc_matches(const DevSegment& S, const MPred* preds, int n_preds, uint32_t doc) {
    bool ok = docset_test(S, S.main, doc);
    if (ok && S.has_deletes) ok = !((S.deleted[doc >> 5] >> (doc & 31)) & 1u);  // searcher.rs:41-46
    for (int k = 0; ok && k < n_preds; k++) {
        const MPred& pr = preds[k];
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
    return ok;
}
#endif

// host: the plan's leading [filter_agg | post_filter_agg_*]* chain -> predicate list; *node = first node below it.
// Returns false when the chain is longer than NARROW_MAXPRED.
static inline bool narrow_chain(const ExecState& es, MPred* preds, int32_t* n_preds, uint32_t* node_out) {
    const PlanMeta& m = *es.meta;
    const uint32_t n_nodes = (uint32_t)m.nodes.size();
    uint32_t node = 0;
    *n_preds = 0;
    while (node < n_nodes && (m.nodes[node].op == TAGG_OP_FILTER || m.nodes[node].op == TAGG_OP_POST_FILTER)) {
        const tagg_node& nd = m.nodes[node];
        if (*n_preds >= NARROW_MAXPRED) return false;
        MPred& pr = preds[(*n_preds)++];
        memset(&pr, 0, sizeof(pr));
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
    *node_out = node;
    return true;
}

// stream_inst_ct_terms.cu — k_stream instantiations: TERMS over one value column with compile-time op masks (C2 / C5 shapes)
#include "stream_kernel.cuh"

template <int BUCKET, int BOPS>
static stream_fn pick_ct_bucket(bool compact, bool stab, bool filt) {
    if (stab && filt) return compact ? (stream_fn)k_stream<Shp<BUCKET, 1, 0, true, true, BOPS, -1, 1>> : (stream_fn)k_stream<Shp<BUCKET, 1, 0, false, true, BOPS, -1, 1>>;
    if (stab) return compact ? (stream_fn)k_stream<Shp<BUCKET, 1, 0, true, true, BOPS, -1, 0>> : (stream_fn)k_stream<Shp<BUCKET, 1, 0, false, true, BOPS, -1, 0>>;
    return compact ? (stream_fn)k_stream<Shp<BUCKET, 1, 0, true, false, BOPS, -1, 0>> : (stream_fn)k_stream<Shp<BUCKET, 1, 0, false, false, BOPS, -1, 0>>;
}
stream_fn stream_pick_ct_terms(uint32_t bops0, bool compact, bool stab, bool filt) {
    switch (bops0) {
        case OPB_MIN: return pick_ct_bucket<BK_TERMS, OPB_MIN>(compact, stab, filt);
        case OPB_MAX: return pick_ct_bucket<BK_TERMS, OPB_MAX>(compact, stab, filt);
        case OPB_SUM: return pick_ct_bucket<BK_TERMS, OPB_SUM>(compact, stab, filt);
        case OPB_MIN | OPB_MAX: return pick_ct_bucket<BK_TERMS, (OPB_MIN | OPB_MAX)>(compact, stab, filt);
        case OPB_SUM | OPB_MIN: return pick_ct_bucket<BK_TERMS, (OPB_SUM | OPB_MIN)>(compact, stab, filt);
        case OPB_SUM | OPB_MAX: return pick_ct_bucket<BK_TERMS, (OPB_SUM | OPB_MAX)>(compact, stab, filt);
        case OPB_MIN | OPB_MAX | OPB_SUM: return pick_ct_bucket<BK_TERMS, (OPB_MIN | OPB_MAX | OPB_SUM)>(compact, stab, filt);
    }
    return nullptr;
}

// stream_inst_ct_root.cu — k_stream instantiations: root metrics over one f64 column with compile-time op masks (C1 shape)
#include "stream_kernel.cuh"

template <int ROPS>
static stream_fn pick_ct_root(bool compact) {
    return compact ? (stream_fn)k_stream<Shp<BK_NONE, 0, 1, true, false, -1, ROPS>> : (stream_fn)k_stream<Shp<BK_NONE, 0, 1, false, false, -1, ROPS>>;
}
stream_fn stream_pick_ct_root(uint32_t rops0, bool compact) {
    switch (rops0) {
        case OPB_MIN: return pick_ct_root<OPB_MIN>(compact);
        case OPB_MAX: return pick_ct_root<OPB_MAX>(compact);
        case OPB_SUM: return pick_ct_root<OPB_SUM>(compact);
        case OPB_MIN | OPB_MAX: return pick_ct_root<(OPB_MIN | OPB_MAX)>(compact);
        case OPB_MIN | OPB_SUM: return pick_ct_root<(OPB_MIN | OPB_SUM)>(compact);
        case OPB_MAX | OPB_SUM: return pick_ct_root<(OPB_MAX | OPB_SUM)>(compact);
        case OPB_MIN | OPB_MAX | OPB_SUM: return pick_ct_root<(OPB_MIN | OPB_MAX | OPB_SUM)>(compact);
    }
    return nullptr;
}

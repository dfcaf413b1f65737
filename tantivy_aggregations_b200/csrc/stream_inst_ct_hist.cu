// stream_inst_ct_hist.cu — k_stream instantiations: count-only histogram and the percentile rank bins (C3 shapes)
#include "stream_kernel.cuh"

stream_fn stream_pick_ct_hist_rank(int bucket, bool rank_linear, bool compact, bool stab) {
    if (bucket == BK_HIST) {
        if (stab) return compact ? (stream_fn)k_stream<Shp<BK_HIST, 0, 0, true, true, 0, -1, 0>> : (stream_fn)k_stream<Shp<BK_HIST, 0, 0, false, true, 0, -1, 0>>;
        return compact ? (stream_fn)k_stream<Shp<BK_HIST, 0, 0, true, false, 0, -1, 0>> : (stream_fn)k_stream<Shp<BK_HIST, 0, 0, false, false, 0, -1, 0>>;
    }
    if (rank_linear) return compact ? (stream_fn)k_stream<Shp<BK_RANK, 1, 0, true, true, (OPB_MIN | OPB_MAX), -2, 1>> : (stream_fn)k_stream<Shp<BK_RANK, 1, 0, false, true, (OPB_MIN | OPB_MAX), -2, 1>>;
    return compact ? (stream_fn)k_stream<Shp<BK_RANK, 1, 0, true, true, (OPB_MIN | OPB_MAX), -1, 1>> : (stream_fn)k_stream<Shp<BK_RANK, 1, 0, false, true, (OPB_MIN | OPB_MAX), -1, 1>>;
}

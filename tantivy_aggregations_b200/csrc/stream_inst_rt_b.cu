// stream_inst_rt_b.cu — k_stream instantiations: runtime-op-mask shapes with a HISTOGRAM node
#include "stream_kernel.cuh"

// runtime-op-mask shapes: any flat plan
template <int BUCKET, int NBG, int NRG>
static stream_fn pick_flags(bool compact, bool stab) {
    if (BUCKET == BK_NONE) stab = false;
    if (stab) return compact ? (stream_fn)k_stream<Shp<BUCKET, NBG, NRG, true, (BUCKET != BK_NONE)>> : (stream_fn)k_stream<Shp<BUCKET, NBG, NRG, false, (BUCKET != BK_NONE)>>;
    return compact ? (stream_fn)k_stream<Shp<BUCKET, NBG, NRG, true, false>> : (stream_fn)k_stream<Shp<BUCKET, NBG, NRG, false, false>>;
}
template <int BUCKET, int NBG>
static stream_fn pick_nrg(int nrg, bool compact, bool stab) {
    switch (nrg) {
        case 0: return pick_flags<BUCKET, NBG, 0>(compact, stab);
        default: return pick_flags<BUCKET, NBG, ST_MAXRG>(compact, stab);
    }
}
template <int BUCKET>
static stream_fn pick_nbg(int nbg, int nrg, bool compact, bool stab) {
    switch (nbg) {
        case 0: return pick_nrg<BUCKET, 0>(nrg, compact, stab);
        default: return pick_nrg<BUCKET, 3>(nrg, compact, stab);
    }
}

stream_fn stream_pick_rt_hist(int nbg, int nrg, bool compact, bool stab) { return pick_nbg<BK_HIST>(nbg, nrg, compact, stab); }

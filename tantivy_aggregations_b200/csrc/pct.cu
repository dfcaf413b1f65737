// pct.cu — K4: percentiles_agg_f64 (percentile.rs:87-90, 163-177) on the streaming path.
//
// The reference inserts every matched value into a CKMS(eps = 0.01) sketch and answers percentile(q) with an element
// whose rank is within +-eps*q*n of q*n.  Here the fruit is a list of EXACT order statistics (rank, value) dense
// enough that every target rank has a stored neighbour well inside that band, produced in ONE pass over the column:
//   1. k_pct_sample gathers ~64k matched values (a systematic sample of the doc stream); the host sorts them and
//      picks two code thresholds lo < hi and a bin width 2^shift such that, going by the sample, every bin in
//      [lo, hi) will hold fewer than eps/2 of the values below it;
//   2. the streaming kernel (stream.cu, BK_RANK) keeps count / min / max per bin in shared memory — the minimum of
//      a bin is the exact order statistic of rank (values below the bin) + 1, its maximum that of rank
//      (values below) + count — and appends the few values outside [lo, hi) to an exact list;
//   3. pct_rank_collect sorts the list, checks the precision bound on the real counts (count <= eps * values below,
//      or a single distinct value) and emits the pairs.  If the check fails (a distribution the equal-width bins
//      cannot resolve) the query is redone on the exact path (materialise + radix sort, generic.cu / result.cu).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <cub/device/device_radix_sort.cuh>

#include "exec.h"
#include "narrow.cuh"

#define PCT_SAMPLE 65536u
#define PCT_BINS 4096u
#define PCT_MIN_DOCS (8ull << 20)  // below this the exact path is cheap enough
#define PCT_EPS 0.01               // percentile.rs:174

struct SampleParams {
    const DevSegment* segs;
    const uint64_t* doc_begin;  // n_segs + 1
    uint32_t n_segs, n_samples;
    uint64_t n_docs;
    int32_t col, n_preds;
    MPred preds[NARROW_MAXPRED];
    uint64_t* out;
    unsigned int* count;
};

__global__ void k_pct_sample(const SampleParams p) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n_samples) return;
    // systematic sample: one document per stratum, at a hashed position inside it
    const uint64_t stratum = p.n_docs / p.n_samples;
    const uint64_t g = (uint64_t)s * stratum + mix64(s + 0x1234567ull) % stratum;
    uint32_t lo = 0, hi = p.n_segs;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (p.doc_begin[mid] <= g) lo = mid; else hi = mid;
    }
    const DevSegment& S = p.segs[lo];
    const uint32_t doc = (uint32_t)(g - p.doc_begin[lo]);
    if (doc >= S.max_doc || !doc_matches(S, p.preds, p.n_preds, doc)) return;
    const uint64_t code = col_get(S.cols[p.col], doc);
    p.out[atomicAdd(p.count, 1u)] = code;
}

__global__ void k_gather_u64(const uint64_t* __restrict__ src, const uint64_t* __restrict__ pos, uint64_t n, uint64_t* __restrict__ dst) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[pos[i]];
}

static inline uint64_t code_to_f64_bits(uint64_t c) { return (c >> 63) ? (c ^ 0x8000000000000000ull) : ~c; }

void pct_rank_release(ExecState& es) {
    for (auto& r : es.rank) {
        if (r.d_block) cudaFreeAsync(r.d_block, es.st);
        if (r.d_tail) cudaFreeAsync(r.d_tail, es.st);
        if (r.d_sorted) cudaFreeAsync(r.d_sorted, es.st);
        if (r.d_cub) cudaFreeAsync(r.d_cub, es.st);
        if (r.d_pick) cudaFreeAsync(r.d_pick, es.st);
        r = ExecState::RankState();
    }
}

// ---- the exact lists, summarised on the device ---------------------------------------------------------------------
// Ranks kept from an exactly sorted run of n values: every rank up to PICK_DENSE, then geometric with ratio 1 + eps/4.
// Past PICK_DENSE the sequence does not depend on n (only its end is clamped to n), so it is tabulated once.
#define PICK_DENSE 4096u
#define PICK_HIGH 8192u  // the high list is thinned to every (n_high / 4096)-th value: fewer than 8192 of them
static const std::vector<uint64_t>& geometric_schedule() {
    static const std::vector<uint64_t> G = [] {
        std::vector<uint64_t> g;
        uint64_t r = PICK_DENSE;
        while (r < (1ull << 40)) {
            r += std::max<uint64_t>(1, (uint64_t)((double)r * 0.0025));
            g.push_back(r);
        }
        return g;
    }();
    return G;
}
static inline size_t pick_words(uint32_t sched_len) { return (size_t)PICK_DENSE + sched_len + PICK_HIGH + 1; }

// pick layout: [PICK_DENSE lowest][sched_len geometric ranks of the low list][PICK_HIGH thinned high list][the largest]
__global__ void k_tail_pick(const uint64_t* __restrict__ sorted, const unsigned long long* __restrict__ tail_count, const uint64_t* __restrict__ sched,
                            uint32_t sched_len, uint64_t n_sorted, uint64_t* __restrict__ pick) {
    const uint64_t n_tail = tail_count[0], n_low = tail_count[1];
    if (n_tail > n_sorted || n_low > n_tail) return;  // longer than predicted: the host path sorts again
    const uint64_t n_high = n_tail - n_low;
    const uint64_t step = n_high / 4096 > 1 ? n_high / 4096 : 1;
    const uint32_t total = PICK_DENSE + sched_len + PICK_HIGH + 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint64_t v = 0;
        if (i < PICK_DENSE) {
            if (i < n_low) v = sorted[i];
        } else if (i < PICK_DENSE + sched_len) {
            const uint32_t j = i - PICK_DENSE;
            const uint64_t prev = j ? sched[j - 1] : PICK_DENSE;
            if (prev < n_low) v = sorted[(sched[j] < n_low ? sched[j] : n_low) - 1];
        } else if (i < PICK_DENSE + sched_len + PICK_HIGH) {
            const uint64_t at = (uint64_t)(i - PICK_DENSE - sched_len) * step;
            if (at < n_high) v = sorted[n_low + at];
        } else if (n_tail) {
            v = sorted[n_tail - 1];
        }
        pick[i] = v;
    }
}

static int sched_on_device(tagg_ctx* ctx) {
    std::lock_guard<std::mutex> g(ctx->mu);
    if (ctx->pct_sched_dev) return 0;
    const std::vector<uint64_t>& G = geometric_schedule();
    uint64_t* d = nullptr;
    if (cudaMalloc((void**)&d, G.size() * 8) != cudaSuccess) { cudaGetLastError(); return 1; }
    if (cudaMemcpy(d, G.data(), G.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); cudaFree(d); return 1; }
    ctx->pct_sched_dev = d;
    ctx->pct_sched_len = (uint32_t)G.size();
    return 0;
}

int pct_rank_plan(ExecState& es, uint32_t node, int k) {
    const PlanMeta& m = *es.meta;
    if (es.no_rank || es.collective || k < 0 || k >= 4) return 0;
    const size_t nseg = es.hsegs.size();
    std::vector<uint64_t> begin(nseg + 1, 0);
    for (size_t i = 0; i < nseg; i++) begin[i + 1] = begin[i] + es.hsegs[i].max_doc;
    const uint64_t n_docs = begin[nseg];
    if (n_docs < PCT_MIN_DOCS) return 0;

    {   // thresholds of an earlier call on the same (plan, segment set): no sample pass
        std::lock_guard<std::mutex> g(es.plan->mu);
        const PctThresholds& pc = es.plan->pct_cache[k];
        if (pc.valid && pc.node == node && pc.segs.size() == nseg && std::equal(pc.segs.begin(), pc.segs.end(), es.segs.begin(), [](const void* a, const tagg_segment* b) { return a == (const void*)b; })) {
            ExecState::RankState& R = es.rank[k];
            R = ExecState::RankState();
            R.lo = pc.lo; R.span = pc.span; R.shift = pc.shift; R.mul = pc.mul; R.n_bins = pc.n_bins;
            R.linear = pc.linear; R.f_lo = pc.f_lo; R.f_scale = pc.f_scale;
            R.from_cache = true;
            R.node = node;
            R.tail_cap = pc.tail_seen * 2 + (1u << 18);
            const size_t blk = 16 + (size_t)R.n_bins * 25;
            if (cudaMallocAsync((void**)&R.d_block, blk, es.st) == cudaSuccess && cudaMallocAsync((void**)&R.d_tail, R.tail_cap * 8 + 16, es.st) == cudaSuccess &&
                cudaMemsetAsync(R.d_block, 0, blk, es.st) == cudaSuccess) {
                R.d_tail_count = (unsigned long long*)R.d_block;
                R.d_count = (uint64_t*)(R.d_block + 16);
                R.d_min = R.d_count + R.n_bins;
                R.d_max = R.d_min + R.n_bins;
                R.d_present = (uint8_t*)(R.d_max + R.n_bins);
                R.active = true;
                // the list length of the earlier call predicts this one's: sort and thin the lists behind the pass
                const uint64_t pred = std::min<uint64_t>(R.tail_cap, pc.tail_seen + pc.tail_seen / 16 + 4096);
                static const bool no_pick = getenv("TAGG_PCT_HOST_LISTS") != nullptr;  // experiment switch
                if (!no_pick && pred < (1ull << 31) && sched_on_device(es.ctx) == 0) {
                    size_t cub_bytes = 0;
                    cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)pred, 0, 64, es.st);
                    if (cudaMallocAsync((void**)&R.d_sorted, pred * 8 + 16, es.st) == cudaSuccess && cudaMallocAsync(&R.d_cub, cub_bytes + 16, es.st) == cudaSuccess &&
                        cudaMallocAsync((void**)&R.d_pick, pick_words(es.ctx->pct_sched_len) * 8, es.st) == cudaSuccess &&
                        cudaMemsetAsync(R.d_tail, 0xff, pred * 8, es.st) == cudaSuccess) {
                        R.tail_pred = pred;
                        R.cub_bytes = cub_bytes;
                    } else {
                        cudaGetLastError();
                        if (R.d_sorted) cudaFreeAsync(R.d_sorted, es.st);
                        if (R.d_cub) cudaFreeAsync(R.d_cub, es.st);
                        if (R.d_pick) cudaFreeAsync(R.d_pick, es.st);
                        R.d_sorted = nullptr; R.d_cub = nullptr; R.d_pick = nullptr;
                    }
                }
                return 1;
            }
            cudaGetLastError();
            pct_rank_release(es);
        }
    }

    SampleParams sp;
    memset(&sp, 0, sizeof(sp));
    uint32_t first = 0;
    if (!narrow_chain(es, sp.preds, &sp.n_preds, &first)) return 0;
    sp.segs = es.d_segs;
    sp.n_segs = (uint32_t)nseg;
    sp.n_samples = PCT_SAMPLE;
    sp.n_docs = n_docs;
    sp.col = m.col_slot[node];

    // host docsets must have landed before the sample looks at them
    if (!es.uploads.empty())
        for (uint32_t c = 0; c < es.n_chunks; c++)
            if (cudaStreamWaitEvent(es.st, es.call->chunk_ev[c], 0) != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "stream ordering failed");
    uint8_t* d_tmp = nullptr;  // [doc_begin][count 16 B][sample][sorted sample][cub scratch]
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)PCT_SAMPLE, 0, 64, es.st);
    const size_t off_begin = 0, off_count = (nseg + 1) * 8, off_sample = off_count + 16, off_sorted = off_sample + PCT_SAMPLE * 8,
                 off_cub = off_sorted + PCT_SAMPLE * 8, total = off_cub + cub_bytes + 16;
    if (cudaMallocAsync((void**)&d_tmp, total, es.st) != cudaSuccess) return -tagg_fail(TAGG_ERR_OOM, "percentile sample allocation failed");
    es.temps.push_back(d_tmp);
    if (cudaMemcpyAsync(d_tmp + off_begin, es.pin(begin.data(), (nseg + 1) * 8), (nseg + 1) * 8, cudaMemcpyHostToDevice, es.st) != cudaSuccess ||
        cudaMemsetAsync(d_tmp + off_count, 0, 16, es.st) != cudaSuccess ||
        cudaMemsetAsync(d_tmp + off_sample, 0xff, PCT_SAMPLE * 8, es.st) != cudaSuccess)  // empty slots sort to the end
        return -tagg_fail(TAGG_ERR_CUDA, "percentile sample setup failed");
    sp.doc_begin = (const uint64_t*)(d_tmp + off_begin);
    sp.count = (unsigned int*)(d_tmp + off_count);
    sp.out = (uint64_t*)(d_tmp + off_sample);
    k_pct_sample<<<PCT_SAMPLE / 256, 256, 0, es.st>>>(sp);
    es.ctx->launches++;
    es.n_launches++;
    if (cub::DeviceRadixSort::SortKeys(d_tmp + off_cub, cub_bytes, (const uint64_t*)(d_tmp + off_sample), (uint64_t*)(d_tmp + off_sorted),
                                       (int)PCT_SAMPLE, 0, 64, es.st) != cudaSuccess)
        return -tagg_fail(TAGG_ERR_CUDA, "percentile sample sort failed");
    std::vector<uint64_t> smp(PCT_SAMPLE);
    unsigned int mcount = 0;
    if (cudaMemcpyAsync(smp.data(), d_tmp + off_sorted, PCT_SAMPLE * 8, cudaMemcpyDeviceToHost, es.st) != cudaSuccess ||
        cudaMemcpyAsync(&mcount, d_tmp + off_count, 4, cudaMemcpyDeviceToHost, es.st) != cudaSuccess ||
        cudaStreamSynchronize(es.st) != cudaSuccess)
        return -tagg_fail(TAGG_ERR_CUDA, "percentile sample download failed: %s", cudaGetErrorString(cudaGetLastError()));
    const uint64_t mm = std::min<uint64_t>(mcount, PCT_SAMPLE);
    if (mm < 8192) return 0;  // a selective query: few values, the exact path sorts them cheaply
    const double n_est = (double)mm / PCT_SAMPLE * (double)n_docs;

    // thresholds: hi at the 99.8 % sample quantile; lo = the smallest sample quantile from which on 16 bins
    // hold fewer than 16 * eps/2 of the values below them
    // (the top 0.2 % go to the exact list as well: a long right tail would otherwise eat the bins)
    const uint64_t j_hi = mm - std::max<uint64_t>(64, mm / 512);
    const uint64_t top = smp[j_hi];
    const uint64_t hi = top == ~0ull ? top : top + 1;
    // lo = the smallest sample quantile for which, going by the sample, every window of 16 bins holds fewer than
    // 16 * eps/2 of the values below it (bins with a single distinct value cost no precision and do not count)
    // Two monotone binnings are tried: equal-width in CODE space (f64 codes are log-like in the value: right for positive,
    // heavy-tailed data such as prices) and, if that cannot meet the bound, equal-width in VALUE space (right for signed or
    // symmetric data, whose codes around zero span every exponent)
    uint64_t lo = 0, span = 0;
    uint32_t shift = 0, mul = 0, n_bins = 0;
    double f_lo = 0.0, f_scale = 0.0;
    bool linear = false;
    uint64_t j_lo = 0;
    bool found = false;
    for (int mode = 0; mode < 2 && !found; mode++) {
        linear = mode == 1;
        const double v_top = code_to_f64_h(hi - 1);
        if (linear && !std::isfinite(v_top)) break;
        for (uint64_t j = std::max<uint64_t>(64, mm / 1024); j <= mm / 8 && !found; j += std::max<uint64_t>(1, j / 4)) {
            lo = smp[j];
            if (hi <= lo) break;
            span = hi - lo;
            {   // 32-bit rank word of a code inside [lo, hi): (code - lo) >> shift (also the kernel's min / max filter word)
                const uint32_t bits = 64 - (uint32_t)__builtin_clzll(span - 1 ? span - 1 : 1);
                shift = bits > 32 ? bits - 32 : 0;
            }
            if (linear) {
                f_lo = code_to_f64_h(lo);
                if (!std::isfinite(f_lo) || !(v_top > f_lo)) break;
                f_scale = (double)PCT_BINS / (v_top - f_lo);
                if (!std::isfinite(f_scale)) break;
                n_bins = PCT_BINS;
            } else {
                // bin = umulhi(d >> shift, mul) (or d itself when the span is at most PCT_BINS codes)
                const uint64_t xmax = (span - 1) >> shift;
                if (xmax + 1 <= PCT_BINS) { mul = 0; n_bins = (uint32_t)(xmax + 1); }
                else { mul = (uint32_t)(((uint64_t)PCT_BINS << 32) / (xmax + 1)); n_bins = (uint32_t)((xmax * mul) >> 32) + 1; }
            }
            auto bin_of = [&](uint64_t code) -> uint32_t {
                if (linear) {
                    const double t = (code_to_f64_h(code) - f_lo) * f_scale;
                    const uint32_t b = t > 0.0 ? (t < 4294967040.0 ? (uint32_t)t : 0xffffff00u) : 0u;
                    return b < PCT_BINS ? b : PCT_BINS - 1;
                }
                const uint64_t x = (code - lo) >> shift;
                return mul ? (uint32_t)((x * mul) >> 32) : (uint32_t)x;
            };
            // window by window (16 bins): a binary search finds the window's end in the sorted sample; only a window that
            // looks too full is inspected bin by bin, to discount bins that hold a single distinct value
            bool ok = true;
            uint64_t below = j;
            for (uint64_t a = j; a <= j_hi && ok;) {
                const uint32_t win = bin_of(smp[a]) >> 4;
                const uint64_t e = (uint64_t)(std::partition_point(smp.begin() + a, smp.begin() + j_hi + 1,
                                                                   [&](uint64_t code) { return (bin_of(code) >> 4) <= win; }) - smp.begin());
                const double allowed = 16.0 * (PCT_EPS / 2) * (double)below;
                if ((double)(e - a) > allowed) {
                    uint64_t spread = 0;
                    for (uint64_t b0 = a; b0 < e;) {
                        const uint32_t bin = bin_of(smp[b0]);
                        uint64_t b1 = b0;
                        while (b1 < e && bin_of(smp[b1]) == bin) b1++;
                        if (b1 - b0 < 2 || smp[b1 - 1] != smp[b0]) spread += b1 - b0;
                        b0 = b1;
                    }
                    if ((double)spread > allowed) ok = false;
                }
                below += e - a;
                a = e;
            }
            if (ok) { found = true; j_lo = j; }
        }
    }
    if (!found) return 0;

    ExecState::RankState& R = es.rank[k];
    R = ExecState::RankState();
    R.lo = lo; R.span = span; R.shift = shift; R.mul = mul; R.n_bins = n_bins;
    R.linear = linear; R.f_lo = f_lo; R.f_scale = f_scale;
    R.node = node;
    // values outside [lo, hi): the sampled share below lo, plus what lies above the largest sampled value
    R.tail_cap = (uint64_t)(n_est * ((double)(j_lo + (mm - j_hi)) / (double)mm) * 1.5) + (uint64_t)(n_est / mm * 64) + (1u << 18);
    const size_t blk = 16 + (size_t)n_bins * 25;
    if (cudaMallocAsync((void**)&R.d_block, blk, es.st) != cudaSuccess || cudaMallocAsync((void**)&R.d_tail, R.tail_cap * 8 + 16, es.st) != cudaSuccess) {
        pct_rank_release(es);
        cudaGetLastError();
        return 0;  // no room for the list: exact path (which reports its own OOM if it must)
    }
    if (cudaMemsetAsync(R.d_block, 0, blk, es.st) != cudaSuccess) return -tagg_fail(TAGG_ERR_CUDA, "rank table clear failed");
    R.d_tail_count = (unsigned long long*)R.d_block;
    R.d_count = (uint64_t*)(R.d_block + 16);
    R.d_min = R.d_count + n_bins;
    R.d_max = R.d_min + n_bins;
    R.d_present = (uint8_t*)(R.d_max + n_bins);
    R.active = true;
    return 1;
}

// Ranks kept from an exactly sorted run of n values: every rank up to 4096, then geometric with ratio 1 + eps/4
static void rank_schedule(uint64_t n, std::vector<uint64_t>& ranks) {
    ranks.clear();
    uint64_t dense = std::min<uint64_t>(n, 4096);
    for (uint64_t r = 1; r <= dense; r++) ranks.push_back(r);
    uint64_t r = dense;
    while (r < n) {
        uint64_t nx = r + std::max<uint64_t>(1, (uint64_t)((double)r * 0.0025));
        if (nx > n) nx = n;
        ranks.push_back(nx);
        r = nx;
    }
}

// the bin tables (count / min / max per bin + the list counters) start their way to pinned host memory right behind the
// pass, so that the pass's own synchronisation delivers them
int pct_rank_prefetch(ExecState& es) {
    for (auto& R : es.rank) {
        if (!R.active || !R.d_block) continue;
        const size_t bytes = 16 + (size_t)R.n_bins * 24;
        uint8_t* h = (uint8_t*)const_cast<void*>(es.pin(nullptr, bytes));
        R.h_block = nullptr;
        if (!h) continue;
        CUDA_TRY(cudaMemcpyAsync(h, R.d_block, bytes, cudaMemcpyDeviceToHost, es.st));
        R.h_block = h;
        R.h_pick = nullptr;
        if (R.tail_pred) {
            const size_t pbytes = pick_words(es.ctx->pct_sched_len) * 8;
            uint64_t* hp = (uint64_t*)const_cast<void*>(es.pin(nullptr, pbytes));
            if (!hp) continue;
            CUDA_TRY(cub::DeviceRadixSort::SortKeys(R.d_cub, R.cub_bytes, (const uint64_t*)R.d_tail, R.d_sorted, (int)R.tail_pred, 0, 64, es.st));
            k_tail_pick<<<32, 256, 0, es.st>>>(R.d_sorted, R.d_tail_count, es.ctx->pct_sched_dev, es.ctx->pct_sched_len, R.tail_pred, R.d_pick);
            es.ctx->launches++;
            es.n_launches++;
            CUDA_TRY(cudaMemcpyAsync(hp, R.d_pick, pbytes, cudaMemcpyDeviceToHost, es.st));
            R.h_pick = hp;
        }
    }
    return 0;
}

int pct_rank_collect(ExecState& es, int k) {
    ExecState::RankState& R = es.rank[k];
    const uint32_t nb = R.n_bins;
    std::vector<uint8_t> blk_copy;
    const uint8_t* blk = R.h_block;
    if (!blk) {
        blk_copy.resize(16 + (size_t)nb * 24);
        CUDA_TRY(cudaMemcpyAsync(blk_copy.data(), R.d_block, blk_copy.size(), cudaMemcpyDeviceToHost, es.st));
        CUDA_TRY(cudaStreamSynchronize(es.st));
        blk = blk_copy.data();
    }
    unsigned long long tc[2];
    memcpy(tc, blk, 16);
    const uint64_t* cnt = (const uint64_t*)(blk + 16);
    const uint64_t* mn = cnt + nb;  // max-form: ~code
    const uint64_t* mx = mn + nb;
    const uint64_t n_tail = tc[0], n_low = tc[1];
    static const bool trace = getenv("TAGG_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[tagg] rank bins (%s): lo=%016llx span=%016llx shift=%u bins=%u tail=%llu (low %llu) cap=%llu\n", R.linear ? "value space" : "code space", (unsigned long long)R.lo,
                       (unsigned long long)R.span, R.shift, nb, (unsigned long long)n_tail, (unsigned long long)n_low, (unsigned long long)R.tail_cap);
    if (n_tail > R.tail_cap || n_low > n_tail) return 0;
    const uint64_t n_high = n_tail - n_low;

    // precision check on the real counts
    uint64_t below = n_low, n_binned = 0;
    for (uint32_t b = 0; b < nb; b++) {
        const uint64_t c = cnt[b];
        if (!c) continue;
        if (c > 1 && ~mn[b] != mx[b] && (double)c > std::max(1.0, PCT_EPS * (double)below)) {
            if (trace) fprintf(stderr, "[tagg] rank bins: bin %u holds %llu values over %llu below it: exact path\n", b, (unsigned long long)c, (unsigned long long)below);
            return 0;
        }
        below += c;
        n_binned += c;
    }
    PctSummary& S = R.summary;
    S = PctSummary();
    S.n_total = n_low + n_binned + n_high;

    // the exact lists: sort, keep every low rank up to 4096 then a geometric schedule; the high end evenly thinned
    std::vector<uint64_t> pos, pos_rank;
    if (n_tail && R.h_pick && n_tail <= R.tail_pred) {
        // the lists were sorted and thinned on the device (k_tail_pick): the same ranks as below, already on the host
        const std::vector<uint64_t>& G = geometric_schedule();
        const uint32_t L = es.ctx->pct_sched_len;
        const uint64_t* pk = R.h_pick;
        for (uint64_t r = 1; r <= std::min<uint64_t>(n_low, PICK_DENSE); r++) { S.ranks.push_back(r); S.value_bits.push_back(code_to_f64_bits(pk[r - 1])); }
        for (uint32_t j = 0; j < L; j++) {
            const uint64_t prev = j ? G[j - 1] : PICK_DENSE;
            if (prev >= n_low) break;
            S.ranks.push_back(std::min<uint64_t>(G[j], n_low));
            S.value_bits.push_back(code_to_f64_bits(pk[PICK_DENSE + j]));
        }
        below = n_low;
        for (uint32_t b = 0; b < nb; b++) {
            if (!cnt[b]) continue;
            S.ranks.push_back(below + 1); S.value_bits.push_back(code_to_f64_bits(~mn[b]));
            if (cnt[b] > 1) { S.ranks.push_back(below + cnt[b]); S.value_bits.push_back(code_to_f64_bits(mx[b])); }
            below += cnt[b];
        }
        const uint64_t step = std::max<uint64_t>(1, n_high / 4096);
        uint64_t last_at = 0, i = 0;
        for (uint64_t at = 0; at < n_high; at += step, i++) {
            S.ranks.push_back(n_low + n_binned + at + 1);
            S.value_bits.push_back(code_to_f64_bits(pk[PICK_DENSE + L + i]));
            last_at = at;
        }
        if (n_high && last_at != n_high - 1) { S.ranks.push_back(S.n_total); S.value_bits.push_back(code_to_f64_bits(pk[PICK_DENSE + L + PICK_HIGH])); }
    } else if (n_tail) {
        uint64_t* d_sorted = nullptr;
        void* d_cub = nullptr;
        size_t cub_bytes = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)n_tail, 0, 64, es.st);
        CUDA_TRY(cudaMallocAsync((void**)&d_sorted, n_tail * 8, es.st));
        es.temps.push_back(d_sorted);
        CUDA_TRY(cudaMallocAsync(&d_cub, cub_bytes ? cub_bytes : 16, es.st));
        es.temps.push_back(d_cub);
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(d_cub, cub_bytes, (const uint64_t*)R.d_tail, d_sorted, (int)n_tail, 0, 64, es.st));
        std::vector<uint64_t> rk;
        rank_schedule(n_low, rk);
        for (uint64_t r : rk) { pos.push_back(r - 1); pos_rank.push_back(r); }
        const uint64_t step = std::max<uint64_t>(1, n_high / 4096);
        for (uint64_t i = 0; i < n_high; i += step) { pos.push_back(n_low + i); pos_rank.push_back(n_low + n_binned + i + 1); }
        if (n_high && pos.back() != n_tail - 1) { pos.push_back(n_tail - 1); pos_rank.push_back(S.n_total); }
        uint64_t *d_pos = nullptr, *d_out = nullptr;
        CUDA_TRY(cudaMallocAsync((void**)&d_pos, pos.size() * 8 + 16, es.st));
        es.temps.push_back(d_pos);
        CUDA_TRY(cudaMallocAsync((void**)&d_out, pos.size() * 8 + 16, es.st));
        es.temps.push_back(d_out);
        std::vector<uint64_t> vals(pos.size());
        if (!pos.empty()) {
            CUDA_TRY(cudaMemcpyAsync(d_pos, pos.data(), pos.size() * 8, cudaMemcpyHostToDevice, es.st));
            k_gather_u64<<<(unsigned)std::min<uint64_t>((pos.size() + 255) / 256, 1024), 256, 0, es.st>>>(d_sorted, d_pos, pos.size(), d_out);
            es.ctx->launches++;
            es.n_launches++;
            CUDA_TRY(cudaMemcpyAsync(vals.data(), d_out, pos.size() * 8, cudaMemcpyDeviceToHost, es.st));
            CUDA_TRY(cudaStreamSynchronize(es.st));
        }
        size_t i = 0;
        for (; i < pos.size() && pos[i] < n_low; i++) { S.ranks.push_back(pos_rank[i]); S.value_bits.push_back(code_to_f64_bits(vals[i])); }
        below = n_low;
        for (uint32_t b = 0; b < nb; b++) {
            if (!cnt[b]) continue;
            S.ranks.push_back(below + 1); S.value_bits.push_back(code_to_f64_bits(~mn[b]));
            if (cnt[b] > 1) { S.ranks.push_back(below + cnt[b]); S.value_bits.push_back(code_to_f64_bits(mx[b])); }
            below += cnt[b];
        }
        for (; i < pos.size(); i++) { S.ranks.push_back(pos_rank[i]); S.value_bits.push_back(code_to_f64_bits(vals[i])); }
    } else {
        below = 0;
        for (uint32_t b = 0; b < nb; b++) {
            if (!cnt[b]) continue;
            S.ranks.push_back(below + 1); S.value_bits.push_back(code_to_f64_bits(~mn[b]));
            if (cnt[b] > 1) { S.ranks.push_back(below + cnt[b]); S.value_bits.push_back(code_to_f64_bits(mx[b])); }
            below += cnt[b];
        }
    }
    {   // remember the thresholds for the next query on this (plan, segment set)
        std::lock_guard<std::mutex> g(es.plan->mu);
        PctThresholds& pc = es.plan->pct_cache[k];
        pc.valid = true;
        pc.segs.assign(es.segs.begin(), es.segs.end());
        pc.node = R.node;
        pc.lo = R.lo; pc.span = R.span; pc.shift = R.shift; pc.mul = R.mul; pc.n_bins = R.n_bins;
        pc.linear = R.linear; pc.f_lo = R.f_lo; pc.f_scale = R.f_scale;
        pc.tail_seen = n_tail;
    }
    return 1;
}

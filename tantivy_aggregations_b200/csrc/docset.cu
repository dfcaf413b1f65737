// docset.cu — K6: docsets as device bitsets.  Producing the matched-doc set on the device (a
// TermQuery / RangeQuery over an INDEXED|FAST field evaluated from the fast-field column, SURVEY
// §8f-1) and keeping reusable filter docsets resident in HBM, so a query does not pay the PCIe
// upload of a bitset per call (SURVEY §7 "PCIe handoff of the docset").
#include <string.h>

#include "host.h"

__global__ void k_range_bitset(DevColumn col, uint64_t n, uint64_t lo, uint64_t hi, uint32_t* __restrict__ words) {
    uint64_t n_round = (n + 31) & ~31ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (uint64_t)gridDim.x * blockDim.x) {
        bool ok = false;
        if (i < n) {
            uint64_t c = col_get(col, i);
            ok = c >= lo && c <= hi;
        }
        uint32_t m = __ballot_sync(0xffffffffu, ok);
        if ((threadIdx.x & 31) == 0) words[i >> 5] = m;
    }
}
__global__ void k_fill_all(uint32_t* __restrict__ words, uint64_t n) {
    uint64_t n_words = (n + 31) / 32;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t base = w * 32;
        words[w] = base + 32 <= n ? 0xffffffffu : ((1u << (uint32_t)(n - base)) - 1u);
    }
}

static size_t bitset_words(uint32_t max_doc) {
    return (((size_t)max_doc + TAGG_TILE_DOCS - 1) / TAGG_TILE_DOCS) * (TAGG_TILE_DOCS / 32) + 16;
}

// evaluates `in` into a freshly allocated device bitset (tile padded, zero tail)
static int docset_to_device(const tagg_segment* seg, const tagg_docset* in, uint32_t** out_words) {
    tagg_ctx* ctx = seg->ctx;
    size_t words = bitset_words(seg->max_doc), need = ((size_t)seg->max_doc + 7) / 8;
    uint32_t* w = nullptr;
    CUDA_TRY(cudaMalloc(&w, words * 4));
    cudaStream_t st = ctx->acquire_stream();
    int rc = 0;
    auto fail = [&](int r) { ctx->release_stream(st); cudaFree(w); return r; };
    cudaError_t e = cudaMemsetAsync(w, 0, words * 4, st);
    if (e != cudaSuccess) return fail(tagg_fail(TAGG_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e)));
    unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)seg->max_doc + 255) / 256 + 1, (uint64_t)ctx->sm_count * 16);
    switch (in->kind) {
        case TAGG_DOCSET_ALL:
            k_fill_all<<<blocks, 256, 0, st>>>(w, seg->max_doc);
            ctx->launches++;
            break;
        case TAGG_DOCSET_BITSET:
            if (need && (!in->data || in->n < need)) return fail(tagg_fail(TAGG_ERR_BAD_ARG, "bitset docset too short"));
            if (need) e = cudaMemcpyAsync(w, in->data, need, cudaMemcpyHostToDevice, st);
            break;
        case TAGG_DOCSET_DEVICE_BITSET:
            if (need) e = cudaMemcpyAsync(w, in->data, need, cudaMemcpyDeviceToDevice, st);
            break;
        case TAGG_DOCSET_SORTED_IDS: {
            if (in->n && !in->data) return fail(tagg_fail(TAGG_ERR_BAD_ARG, "null doc-id list"));
            uint32_t* ids = nullptr;
            if (in->n) {
                uint32_t bad = 0;  // [ids][flag]: an id >= max_doc or a list that is not strictly ascending
                e = cudaMalloc(&ids, in->n * 4 + 16);
                if (e == cudaSuccess) e = cudaMemsetAsync(ids + in->n, 0, 16, st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(ids, in->data, in->n * 4, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = launch_ids_to_bitset(ids, in->n, w, seg->max_doc, ids + in->n, st);
                ctx->launches++;
                if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, ids + in->n, 4, cudaMemcpyDeviceToHost, st);
                cudaStreamSynchronize(st);
                cudaFree(ids);
                if (e == cudaSuccess && bad)
                    return fail(tagg_fail(TAGG_ERR_BAD_ARG, "sorted-id docset: ids must be strictly ascending and < max_doc (%u)", seg->max_doc));
            }
            break;
        }
        case TAGG_DOCSET_COLUMN_RANGE: {
            auto it = seg->cols.find(in->field_id);
            if (it == seg->cols.end())
                return fail(tagg_fail(TAGG_ERR_NO_SUCH_COLUMN, "docset field %u is not a single-valued fast field of the segment", in->field_id));
            if (seg->max_doc) {
                k_range_bitset<<<blocks, 256, 0, st>>>(it->second.dev(), seg->max_doc, in->lo, in->hi, w);
                ctx->launches++;
            }
            break;
        }
        default: return fail(tagg_fail(TAGG_ERR_BAD_ARG, "unknown docset kind %d", in->kind));
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(tagg_fail(TAGG_ERR_CUDA, "docset evaluation failed: %s", cudaGetErrorString(e)));
    ctx->release_stream(st);
    *out_words = w;
    return rc;
}

extern "C" {

int tagg_docset_cache(tagg_segment* seg, const tagg_docset* in, tagg_docset* out) {
    if (!seg || !in || !out) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(seg->ctx->device));
    uint32_t* w = nullptr;
    int rc = docset_to_device(seg, in, &w);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> g(seg->mu);
        seg->cached_bitsets.push_back(w);
    }
    memset(out, 0, sizeof(*out));
    out->kind = TAGG_DOCSET_DEVICE_BITSET;
    out->data = w;
    out->n = ((uint64_t)seg->max_doc + 7) / 8;
    return 0;
}

int tagg_docset_uncache(tagg_segment* seg, const tagg_docset* cached) {
    if (!seg || !cached) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    std::lock_guard<std::mutex> g(seg->mu);
    for (size_t i = 0; i < seg->cached_bitsets.size(); i++)
        if (seg->cached_bitsets[i] == cached->data) {
            cudaSetDevice(seg->ctx->device);
            cudaFree(seg->cached_bitsets[i]);
            seg->cached_bitsets.erase(seg->cached_bitsets.begin() + i);
            return 0;
        }
    return tagg_fail(TAGG_ERR_BAD_ARG, "not a docset cached on this segment");
}

int tagg_docset_to_bitset(const tagg_segment* seg, const tagg_docset* in, uint8_t* out, size_t cap) {
    if (!seg || !in || (!out && seg->max_doc)) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    size_t need = ((size_t)seg->max_doc + 7) / 8;
    if (cap < need) return tagg_fail(TAGG_ERR_BAD_ARG, "buffer too small: %zu < %zu", cap, need);
    CUDA_TRY(cudaSetDevice(seg->ctx->device));
    uint32_t* w = nullptr;
    int rc = docset_to_device(seg, in, &w);
    if (rc) return rc;
    cudaError_t e = need ? cudaMemcpy(out, w, need, cudaMemcpyDeviceToHost) : cudaSuccess;
    cudaFree(w);
    if (e != cudaSuccess) return tagg_fail(TAGG_ERR_CUDA, "bitset download failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"

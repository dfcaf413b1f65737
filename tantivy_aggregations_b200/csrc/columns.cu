// columns.cu — segments and fast-field columns resident in HBM.
//
// Replaces `SegmentReader::fast_fields().{u64,i64,f64,date,u64s,..}(field)` of the reference's
// `for_segment` (sum.rs:49-57, minmax.rs:49-57, terms.rs:75-83, histogram.rs:80-88,
// percentile.rs:48-56, post_filter.rs:196-202).  Columns keep tantivy's bit-packed layout
//  so that real `.fast` bytes can be copied in unchanged; decoded codes are
// re-packed on the device into the same layout (k_pack).
#include <cub/device/device_scan.cuh>

#include <vector>

#include "host.h"

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_minmax(const uint64_t* __restrict__ codes, uint64_t n, unsigned long long* out /* [min,max] */) {
    uint64_t mn = ~0ull, mx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = codes[i];
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(out, (unsigned long long)mn);
        atomicMax(out + 1, (unsigned long long)mx);
    }
}

// tantivy BitPacker restated for parallel writes: output word w gathers every value that
// overlaps bits [64w, 64w+64) of the LSB-first stream.  No atomics: one thread owns one word.
__global__ void k_pack(const uint64_t* __restrict__ codes, uint64_t n, uint64_t min_value, uint32_t nb,
                       uint64_t* __restrict__ out, uint64_t n_words) {
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t bit0 = w * 64;
        uint64_t i = bit0 / nb;
        uint64_t acc = 0;
        for (; i < n; i++) {
            uint64_t vb = i * nb;  // first bit of value i
            if (vb >= bit0 + 64) break;
            uint64_t d = codes[i] - min_value;
            if (vb >= bit0) acc |= d << (vb - bit0);
            else acc |= d >> (bit0 - vb);
        }
        out[w] = acc;
    }
}

// SORTED_IDS docset -> bitset.  The ids come from the caller: an id the segment does not have, or a list that is not
// strictly ascending (what a tantivy scorer yields), is dropped / flagged instead of written out of bounds.
__global__ void k_ids_to_bitset(const uint32_t* __restrict__ ids, uint64_t n, uint32_t* words, uint32_t max_doc, uint32_t* bad) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t d = ids[i];
        if (d >= max_doc || (i > 0 && ids[i - 1] >= d)) {
            if (bad) *bad = 1u;
            if (d >= max_doc) continue;
        }
        atomicOr(words + (d >> 5), 1u << (d & 31));
    }
}

cudaError_t launch_ids_to_bitset(const uint32_t* ids, uint64_t n, uint32_t* words, uint32_t max_doc, uint32_t* bad, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    unsigned blocks = (unsigned)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
    k_ids_to_bitset<<<blocks, 256, 0, stream>>>(ids, n, words, max_doc, bad);
    return cudaGetLastError();
}

// ---- synthetic recipes (include/tagg_synth.h; restated on the CPU in oracle/oracle.cpp) -------------
__device__ __forceinline__ uint64_t synth_x(uint64_t seed, uint64_t tag, uint64_t doc) {
    return mix64(seed ^ tag ^ (doc * 0x9E3779B97F4A7C15ull));
}
__device__ __forceinline__ uint64_t synth_value(int recipe, uint64_t x, uint64_t a, uint64_t b, uint64_t c) {
    if (recipe == 0) {
        // explicit round-to-nearest ops: no FMA contraction, bit-identical to the host recipe
        double u = __dmul_rn((double)(x >> 11), 1.0 / 9007199254740992.0);
        double t = __dmul_rn(100.0, u);
        double v = __dadd_rn(1.0, t);
        return f64_to_code(v);
    }
    if (recipe == 1) return a + x % b;
    if (recipe == 3) {  // power-law (Zipf-like) keys: a + floor(b * u^4), u uniform in [0, 1) — integer arithmetic only
        const uint64_t u = x >> 32, u2 = (u * u) >> 32, u4 = (u2 * u2) >> 32;
        return a + ((u4 * (b & 0xffffffffull)) >> 32);
    }
    return a + (x % b) * c;
}
__global__ void k_synth(int recipe, uint64_t seed, uint64_t tag, uint64_t doc_base, uint64_t n, uint64_t a,
                        uint64_t b, uint64_t c, uint64_t* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = synth_value(recipe, synth_x(seed, tag, doc_base + i), a, b, c);
}
__global__ void k_synth_counts(uint64_t seed, uint64_t tag, uint64_t doc_base, uint64_t n, uint64_t count_mod,
                               uint64_t* __restrict__ counts /* n+1, last = 0 */) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x)
        counts[i] = i < n ? synth_x(seed, tag ^ 0xC0FFEE1234567ull, doc_base + i) % count_mod : 0;
}
__global__ void k_synth_multi(int recipe, uint64_t seed, uint64_t tag, uint64_t doc_base, uint64_t n, uint64_t a,
                              uint64_t b, uint64_t c, const uint64_t* __restrict__ offsets, uint64_t* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = synth_x(seed, tag, doc_base + i);
        uint64_t s = offsets[i], e = offsets[i + 1];
        for (uint64_t j = 0; s + j < e; j++)
            out[s + j] = synth_value(recipe, mix64(x + (j + 1) * 0xD6E8FEB86659FD93ull), a, b, c);
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static inline uint32_t compute_num_bits(uint64_t amplitude) {  // tantivy common::compute_num_bits
    uint32_t b = amplitude == 0 ? 0 : 64 - (uint32_t)__builtin_clzll(amplitude);
    return b <= 56 ? b : 64;
}
static inline unsigned grid_for(uint64_t n, int sm_count) {
    uint64_t blocks = (n + 255) / 256;
    uint64_t cap = (uint64_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    return (unsigned)blocks;
}

static int column_alloc(HostColumn* c) {
    c->payload_bytes = (size_t)((c->n_values * c->num_bits + 7) / 8);
    size_t tile_bytes = (size_t)(TAGG_TILE_DOCS / 8) * (c->num_bits ? c->num_bits : 1);
    size_t tiles = (size_t)((c->n_values + TAGG_TILE_DOCS - 1) / TAGG_TILE_DOCS);
    c->alloc_bytes = (tiles ? tiles : 1) * tile_bytes + 256;
    CUDA_TRY(cudaMalloc(&c->dptr, c->alloc_bytes));
    return 0;
}

void column_free(HostColumn* c) {
    if (c->dptr) cudaFree(c->dptr);
    c->dptr = nullptr;
}

int column_from_bytes(tagg_ctx* ctx, int kind, const uint8_t* bytes, size_t len, uint64_t n_values, HostColumn* out) {
    (void)ctx;
    if (!bytes || len < 16) return tagg_fail(TAGG_ERR_BAD_ARG, "column bytes shorter than the 16-byte header");
    HostColumn c;
    c.kind = kind;
    c.n_values = n_values;
    for (int i = 0; i < 8; i++) c.min_value |= (uint64_t)bytes[i] << (8 * i);
    for (int i = 0; i < 8; i++) c.amplitude |= (uint64_t)bytes[8 + i] << (8 * i);
    c.num_bits = compute_num_bits(c.amplitude);
    int rc = column_alloc(&c);
    if (rc) return rc;
    if (len - 16 < c.payload_bytes) {
        column_free(&c);
        return tagg_fail(TAGG_ERR_BAD_ARG, "column bytes hold %zu payload bytes, %zu needed for %llu values of %u bits",
                         len - 16, c.payload_bytes, (unsigned long long)n_values, c.num_bits);
    }
    CUDA_TRY(cudaMemset(c.dptr, 0, c.alloc_bytes));
    if (c.payload_bytes) CUDA_TRY(cudaMemcpy(c.dptr, bytes + 16, c.payload_bytes, cudaMemcpyHostToDevice));
    *out = c;
    return 0;
}

int column_from_device_codes(tagg_ctx* ctx, int kind, const uint64_t* d_codes, uint64_t n, HostColumn* out,
                             cudaStream_t stream) {
    HostColumn c;
    c.kind = kind;
    c.n_values = n;
    if (n) {
        unsigned long long h_mm[2] = {~0ull, 0ull};
        unsigned long long* d_mm = nullptr;
        CUDA_TRY(cudaMalloc(&d_mm, 16));
        CUDA_TRY(cudaMemcpyAsync(d_mm, h_mm, 16, cudaMemcpyHostToDevice, stream));
        k_minmax<<<grid_for(n, ctx->sm_count), 256, 0, stream>>>(d_codes, n, d_mm);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(h_mm, d_mm, 16, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(d_mm);
        c.min_value = h_mm[0];
        c.amplitude = h_mm[1] - h_mm[0];
    }
    c.num_bits = compute_num_bits(c.amplitude);
    int rc = column_alloc(&c);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(c.dptr, 0, c.alloc_bytes, stream));
    if (n && c.num_bits) {
        uint64_t n_words = (n * c.num_bits + 63) / 64;
        k_pack<<<grid_for(n_words, ctx->sm_count), 256, 0, stream>>>(d_codes, n, c.min_value, c.num_bits,
                                                                      (uint64_t*)c.dptr, n_words);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    *out = c;
    return 0;
}

static int upload_codes_column(tagg_ctx* ctx, int kind, const uint64_t* codes, size_t n, HostColumn* out) {
    uint64_t* d = nullptr;
    cudaStream_t st = ctx->acquire_stream();
    int rc = 0;
    if (n) {
        cudaError_t e = cudaMalloc(&d, n * 8);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d, codes, n * 8, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) {
            ctx->release_stream(st);
            if (d) cudaFree(d);
            return tagg_fail(e == cudaErrorMemoryAllocation ? TAGG_ERR_OOM : TAGG_ERR_CUDA, "upload of %zu codes failed: %s",
                             n, cudaGetErrorString(e));
        }
    }
    rc = column_from_device_codes(ctx, kind, d, n, out, st);
    ctx->release_stream(st);
    if (d) cudaFree(d);
    return rc;
}

static bool valid_kind(int k) { return k >= TAGG_U64 && k <= TAGG_DATE; }

extern "C" {

int tagg_segment_create(tagg_ctx* ctx, uint32_t max_doc, tagg_segment** out) {
    if (!ctx || !out) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_segment_create: null argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    auto* s = new tagg_segment();
    s->ctx = ctx;
    s->max_doc = max_doc;
    *out = s;
    return 0;
}

int tagg_segment_destroy(tagg_segment* seg) {
    if (!seg) return 0;
    cudaSetDevice(seg->ctx->device);
    for (auto& kv : seg->cols) column_free(&kv.second);
    for (auto& kv : seg->mcols) { column_free(&kv.second.first); column_free(&kv.second.second); }
    if (seg->d_deleted) cudaFree(seg->d_deleted);
    for (auto p : seg->cached_bitsets) cudaFree(p);
    delete seg;
    return 0;
}

int tagg_segment_max_doc(const tagg_segment* seg, uint32_t* out) {
    if (!seg || !out) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    *out = seg->max_doc;
    return 0;
}

static void drop_field(tagg_segment* seg, uint32_t field_id) {
    auto it = seg->cols.find(field_id);
    if (it != seg->cols.end()) { column_free(&it->second); seg->cols.erase(it); }
    auto mt = seg->mcols.find(field_id);
    if (mt != seg->mcols.end()) { column_free(&mt->second.first); column_free(&mt->second.second); seg->mcols.erase(mt); }
}

int tagg_column_upload(tagg_segment* seg, uint32_t field_id, int kind, const uint8_t* bytes, size_t len) {
    if (!seg || !valid_kind(kind)) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_column_upload: bad argument");
    CUDA_TRY(cudaSetDevice(seg->ctx->device));
    HostColumn c;
    int rc = column_from_bytes(seg->ctx, kind, bytes, len, seg->max_doc, &c);
    if (rc) return rc;
    drop_field(seg, field_id);
    seg->cols[field_id] = c;
    return 0;
}

// ---- tantivy CompositeFile (`.fast` file of a segment) -------------------------------------------------------------------
// Layout restated from tantivy@14735ce src/common/composite_file.rs (external to the reference, see oracle/oracle.cpp
// header — NOT pinned to real tantivy bytes here: no tantivy in the image):
//   [payload of (field, idx) #0][payload #1]...[footer][footer_len: u32 LE]
//   footer = VInt(n) then n x { VInt(offset delta), field: u32 LE, VInt(idx) }, entries ascending by offset;
//   payload i spans [offset_i, offset_{i+1}) and the last one ends where the footer starts.
//   VInt (src/common/vint.rs): 7 bits per byte, least significant group first, the LAST byte carries the 0x80 stop bit.
// A single-valued fast field is (field, 0); a multi-valued one is (field, 0) = offsets column, (field, 1) = values column.
struct FastEntry { uint32_t field, idx; uint64_t begin, end; };
static bool read_vint(const uint8_t*& p, const uint8_t* end, uint64_t* out) {
    uint64_t v = 0;
    for (int shift = 0; p < end && shift < 64; shift += 7) {
        const uint8_t b = *p++;
        v |= (uint64_t)(b & 0x7f) << shift;
        if (b & 0x80) { *out = v; return true; }
    }
    return false;
}
static int parse_fast_file(const uint8_t* bytes, size_t len, std::vector<FastEntry>& out) {
    if (!bytes || len < 5) return tagg_fail(TAGG_ERR_BAD_ARG, "fast file shorter than its footer");
    uint32_t footer_len;
    memcpy(&footer_len, bytes + len - 4, 4);
    if ((size_t)footer_len + 4 > len) return tagg_fail(TAGG_ERR_BAD_ARG, "fast file: footer length %u exceeds the file", footer_len);
    const size_t footer_start = len - 4 - footer_len;
    const uint8_t *p = bytes + footer_start, *end = bytes + len - 4;
    uint64_t n = 0, offset = 0;
    if (!read_vint(p, end, &n) || n > (1u << 20)) return tagg_fail(TAGG_ERR_BAD_ARG, "fast file: malformed footer");
    out.clear();
    for (uint64_t i = 0; i < n; i++) {
        uint64_t delta = 0, idx = 0;
        uint32_t field;
        if (!read_vint(p, end, &delta) || p + 4 > end) return tagg_fail(TAGG_ERR_BAD_ARG, "fast file: malformed footer entry %llu", (unsigned long long)i);
        memcpy(&field, p, 4);
        p += 4;
        if (!read_vint(p, end, &idx)) return tagg_fail(TAGG_ERR_BAD_ARG, "fast file: malformed footer entry %llu", (unsigned long long)i);
        offset += delta;
        if (offset > footer_start) return tagg_fail(TAGG_ERR_BAD_ARG, "fast file: payload offset beyond the footer");
        if (!out.empty()) out.back().end = offset;
        out.push_back({field, (uint32_t)idx, offset, footer_start});
    }
    return 0;
}

int tagg_fast_file_entries(const uint8_t* bytes, size_t len, uint32_t* fields, uint32_t* idxs, uint64_t* begins, uint64_t* ends,
                           uint32_t cap, uint32_t* n_out) {
    if (!n_out) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    std::vector<FastEntry> e;
    int rc = parse_fast_file(bytes, len, e);
    if (rc) return rc;
    *n_out = (uint32_t)e.size();
    for (uint32_t i = 0; i < e.size() && i < cap; i++) {
        if (fields) fields[i] = e[i].field;
        if (idxs) idxs[i] = e[i].idx;
        if (begins) begins[i] = e[i].begin;
        if (ends) ends[i] = e[i].end;
    }
    return 0;
}

int tagg_segment_load_fast_file(tagg_segment* seg, const uint8_t* bytes, size_t len, const tagg_fast_field* fields, uint32_t n_fields) {
    if (!seg || (n_fields && !fields)) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_segment_load_fast_file: bad argument");
    std::vector<FastEntry> e;
    int rc = parse_fast_file(bytes, len, e);
    if (rc) return rc;
    auto find = [&](uint32_t field, uint32_t idx) -> const FastEntry* {
        for (auto& x : e)
            if (x.field == field && x.idx == idx) return &x;
        return nullptr;
    };
    for (uint32_t i = 0; i < n_fields; i++) {
        const tagg_fast_field& f = fields[i];
        const FastEntry* a = find(f.field_id, 0);
        if (!a) return tagg_fail(TAGG_ERR_NO_SUCH_COLUMN, "field %u is not in the fast file", f.field_id);
        if (f.multi) {
            const FastEntry* b = find(f.field_id, 1);
            if (!b) return tagg_fail(TAGG_ERR_NO_SUCH_COLUMN, "field %u has no values column in the fast file (not multi-valued?)", f.field_id);
            rc = tagg_multicolumn_upload(seg, f.field_id, f.kind, bytes + a->begin, (size_t)(a->end - a->begin), bytes + b->begin, (size_t)(b->end - b->begin));
        } else {
            rc = tagg_column_upload(seg, f.field_id, f.kind, bytes + a->begin, (size_t)(a->end - a->begin));
        }
        if (rc) return rc;
    }
    return 0;
}

int tagg_column_upload_codes(tagg_segment* seg, uint32_t field_id, int kind, const uint64_t* codes, size_t n) {
    if (!seg || !valid_kind(kind) || (n && !codes)) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_column_upload_codes: bad argument");
    if (n != seg->max_doc) return tagg_fail(TAGG_ERR_BAD_ARG, "single-valued column needs max_doc=%u codes, got %zu", seg->max_doc, n);
    CUDA_TRY(cudaSetDevice(seg->ctx->device));
    HostColumn c;
    int rc = upload_codes_column(seg->ctx, kind, codes, n, &c);
    if (rc) return rc;
    drop_field(seg, field_id);
    seg->cols[field_id] = c;
    return 0;
}

int tagg_multicolumn_upload(tagg_segment* seg, uint32_t field_id, int kind, const uint8_t* idx_bytes, size_t idx_len,
                            const uint8_t* vals_bytes, size_t vals_len) {
    if (!seg || !valid_kind(kind)) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_multicolumn_upload: bad argument");
    CUDA_TRY(cudaSetDevice(seg->ctx->device));
    if (!idx_bytes || idx_len < 16) return tagg_fail(TAGG_ERR_BAD_ARG, "idx column too short");
    // total values = last offset = min_value + (last packed delta): read it from the idx bytes
    HostColumn idx, vals;
    int rc = column_from_bytes(seg->ctx, TAGG_U64, idx_bytes, idx_len, (uint64_t)seg->max_doc + 1, &idx);
    if (rc) return rc;
    uint64_t total;
    {
        uint64_t i = seg->max_doc;
        if (idx.num_bits == 0) total = idx.min_value;
        else {
            uint64_t bit = i * idx.num_bits, addr = bit >> 3;
            uint64_t w = 0;
            for (int k = 0; k < 8 && 16 + addr + k < idx_len; k++) w |= (uint64_t)idx_bytes[16 + addr + k] << (8 * k);
            uint64_t mask = idx.num_bits == 64 ? ~0ull : ((1ull << idx.num_bits) - 1);
            total = ((w >> (bit & 7)) & mask) + idx.min_value;
        }
    }
    rc = column_from_bytes(seg->ctx, kind, vals_bytes, vals_len, total, &vals);
    if (rc) { column_free(&idx); return rc; }
    drop_field(seg, field_id);
    seg->mcols[field_id] = std::make_pair(idx, vals);
    return 0;
}

int tagg_multicolumn_upload_codes(tagg_segment* seg, uint32_t field_id, int kind, const uint64_t* offsets, size_t n_offsets,
                                  const uint64_t* codes, size_t n_codes) {
    if (!seg || !valid_kind(kind) || !offsets) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_multicolumn_upload_codes: bad argument");
    if (n_offsets != (size_t)seg->max_doc + 1) return tagg_fail(TAGG_ERR_BAD_ARG, "multi-valued column needs max_doc+1 offsets");
    if (offsets[n_offsets - 1] != n_codes) return tagg_fail(TAGG_ERR_BAD_ARG, "last offset must equal the number of values");
    CUDA_TRY(cudaSetDevice(seg->ctx->device));
    HostColumn idx, vals;
    int rc = upload_codes_column(seg->ctx, TAGG_U64, offsets, n_offsets, &idx);
    if (rc) return rc;
    rc = upload_codes_column(seg->ctx, kind, codes, n_codes, &vals);
    if (rc) { column_free(&idx); return rc; }
    drop_field(seg, field_id);
    seg->mcols[field_id] = std::make_pair(idx, vals);
    return 0;
}

int tagg_segment_set_deletes(tagg_segment* seg, const uint8_t* bytes, size_t len) {
    if (!seg || !bytes) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_segment_set_deletes: null argument");
    size_t need = ((size_t)seg->max_doc + 7) / 8;
    if (len < need) return tagg_fail(TAGG_ERR_BAD_ARG, "delete bitset needs %zu bytes, got %zu", need, len);
    CUDA_TRY(cudaSetDevice(seg->ctx->device));
    size_t words = (((size_t)seg->max_doc + TAGG_TILE_DOCS - 1) / TAGG_TILE_DOCS) * (TAGG_TILE_DOCS / 32) + 16;
    if (seg->d_deleted) { cudaFree(seg->d_deleted); seg->d_deleted = nullptr; }
    CUDA_TRY(cudaMalloc(&seg->d_deleted, words * 4));
    CUDA_TRY(cudaMemset(seg->d_deleted, 0, words * 4));
    CUDA_TRY(cudaMemcpy(seg->d_deleted, bytes, need, cudaMemcpyHostToDevice));
    uint64_t n = 0;
    for (size_t i = 0; i < need; i++) {
        uint8_t b = bytes[i];
        if (i == need - 1 && (seg->max_doc & 7)) b &= (uint8_t)((1u << (seg->max_doc & 7)) - 1);
        n += __builtin_popcount(b);
    }
    seg->n_deleted = n;
    seg->has_deletes = true;
    return 0;
}

static const HostColumn* find_column(const tagg_segment* seg, uint32_t field_id, int which) {
    if (which == 0) {
        auto it = seg->cols.find(field_id);
        if (it != seg->cols.end()) return &it->second;
    }
    auto mt = seg->mcols.find(field_id);
    if (mt != seg->mcols.end()) return which == 1 ? &mt->second.first : &mt->second.second;
    return nullptr;
}

int tagg_column_info(const tagg_segment* seg, uint32_t field_id, int which, uint64_t* min_value, uint64_t* amplitude,
                     uint32_t* num_bits, uint64_t* n_values, uint64_t* packed_len) {
    if (!seg) return tagg_fail(TAGG_ERR_BAD_ARG, "null segment");
    const HostColumn* c = find_column(seg, field_id, which);
    if (!c) return tagg_fail(TAGG_ERR_NO_SUCH_COLUMN, "field %u is not a fast field of this segment", field_id);
    if (min_value) *min_value = c->min_value;
    if (amplitude) *amplitude = c->amplitude;
    if (num_bits) *num_bits = c->num_bits;
    if (n_values) *n_values = c->n_values;
    if (packed_len) *packed_len = 16 + c->payload_bytes + 7;  // tantivy's on-disk length
    return 0;
}

int tagg_column_download(const tagg_segment* seg, uint32_t field_id, int which, uint8_t* out, size_t cap) {
    if (!seg || !out) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    const HostColumn* c = find_column(seg, field_id, which);
    if (!c) return tagg_fail(TAGG_ERR_NO_SUCH_COLUMN, "field %u is not a fast field of this segment", field_id);
    size_t len = 16 + c->payload_bytes + 7;
    if (cap < len) return tagg_fail(TAGG_ERR_BAD_ARG, "buffer too small: %zu < %zu", cap, len);
    CUDA_TRY(cudaSetDevice(seg->ctx->device));
    for (int i = 0; i < 8; i++) out[i] = (uint8_t)(c->min_value >> (8 * i));
    for (int i = 0; i < 8; i++) out[8 + i] = (uint8_t)(c->amplitude >> (8 * i));
    if (c->payload_bytes) CUDA_TRY(cudaMemcpy(out + 16, c->dptr, c->payload_bytes, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 7; i++) out[16 + c->payload_bytes + i] = 0;
    return 0;
}

// ---- document address column (include/tagg.h) ---------------------------------------------------
__global__ void k_iota(uint64_t base, uint64_t n, uint64_t* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = base + i;
}
int tagg_segment_doc_address_column(tagg_segment* seg, uint32_t field_id, uint64_t base) {
    if (!seg) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_segment_doc_address_column: null segment");
    tagg_ctx* ctx = seg->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const uint64_t n = seg->max_doc;
    uint64_t* d = nullptr;
    if (n) CUDA_TRY(cudaMalloc(&d, n * 8));
    cudaStream_t st = ctx->acquire_stream();
    if (n) {
        k_iota<<<grid_for(n, ctx->sm_count), 256, 0, st>>>(base, n, d);
        ctx->launches++;
    }
    HostColumn col;
    int rc = column_from_device_codes(ctx, TAGG_U64, d, n, &col, st);
    ctx->release_stream(st);
    if (d) cudaFree(d);
    if (rc) return rc;
    drop_field(seg, field_id);
    seg->cols[field_id] = col;
    return 0;
}

// ---- synthetic generators ------------------------------------------------------------------------
int tagg_synth_column(tagg_segment* seg, uint32_t field_id, int kind, int recipe, uint64_t seed, uint64_t tag,
                      uint64_t doc_base, uint64_t a, uint64_t b, uint64_t c) {
    if (!seg || !valid_kind(kind) || recipe < 0 || recipe > 3 || (recipe && b == 0))
        return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_synth_column: bad argument");
    tagg_ctx* ctx = seg->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t n = seg->max_doc;
    uint64_t* d = nullptr;
    if (n) CUDA_TRY(cudaMalloc(&d, n * 8));
    cudaStream_t st = ctx->acquire_stream();
    if (n) {
        k_synth<<<grid_for(n, ctx->sm_count), 256, 0, st>>>(recipe, seed, tag, doc_base, n, a, b, c, d);
        ctx->launches++;
    }
    HostColumn col;
    int rc = column_from_device_codes(ctx, kind, d, n, &col, st);
    ctx->release_stream(st);
    if (d) cudaFree(d);
    if (rc) return rc;
    drop_field(seg, field_id);
    seg->cols[field_id] = col;
    return 0;
}

int tagg_synth_multicolumn(tagg_segment* seg, uint32_t field_id, int kind, int recipe, uint64_t seed, uint64_t tag,
                           uint64_t doc_base, uint64_t count_mod, uint64_t a, uint64_t b, uint64_t c) {
    if (!seg || !valid_kind(kind) || recipe < 0 || recipe > 3 || (recipe && b == 0) || count_mod == 0)
        return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_synth_multicolumn: bad argument");
    tagg_ctx* ctx = seg->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    uint64_t n = seg->max_doc;
    uint64_t *d_counts = nullptr, *d_offsets = nullptr, *d_vals = nullptr;
    void* d_tmp = nullptr;
    CUDA_TRY(cudaMalloc(&d_counts, (n + 1) * 8));
    CUDA_TRY(cudaMalloc(&d_offsets, (n + 1) * 8));
    cudaStream_t st = ctx->acquire_stream();
    k_synth_counts<<<grid_for(n + 1, ctx->sm_count), 256, 0, st>>>(seed, tag, doc_base, n, count_mod, d_counts);
    ctx->launches++;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, d_offsets, n + 1, st);
    CUDA_TRY(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 16));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_counts, d_offsets, n + 1, st);
    uint64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, d_offsets + n, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (total) CUDA_TRY(cudaMalloc(&d_vals, total * 8));
    if (total && n) {
        k_synth_multi<<<grid_for(n, ctx->sm_count), 256, 0, st>>>(recipe, seed, tag, doc_base, n, a, b, c, d_offsets, d_vals);
        ctx->launches++;
    }
    HostColumn idx, vals;
    int rc = column_from_device_codes(ctx, TAGG_U64, d_offsets, n + 1, &idx, st);
    if (!rc) rc = column_from_device_codes(ctx, kind, d_vals, total, &vals, st);
    ctx->release_stream(st);
    cudaFree(d_counts); cudaFree(d_offsets); cudaFree(d_tmp);
    if (d_vals) cudaFree(d_vals);
    if (rc) return rc;
    drop_field(seg, field_id);
    seg->mcols[field_id] = std::make_pair(idx, vals);
    return 0;
}

}  // extern "C"

// comm.cu — the ONE exchange step of the path: merging bucket tables across GPUs.
//
// The reference merges per-segment fruits on the host, in segment order
// (`PreparedAgg::merge`, src/searcher.rs:93-96).  Here segments are sharded over GPUs (one process
// per GPU); every rank folds its segments into device accumulators laid out identically on all
// ranks (same absolute key domains, agreed by a tiny min/max all-reduce), and one grouped NCCL
// all-reduce over NVLink merges them in place: counts / integer sums -> sum(u64), f64 sums ->
// sum(f64), min / max -> max(u64) on order-preserving codes (MIN is stored complemented), bucket
// existence and Option flags -> max(u8).
//
// NCCL is loaded with dlopen at tagg_comm_init so single-GPU users carry no NCCL dependency and a
// host process that already loaded a libnccl.so.2 (e.g. through torch) shares that copy.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>

#include "exec.h"

// Small tables: every rank gathers all arenas (one collective) and reduces them itself, class by class, in rank
// order — one NCCL call instead of one per reduction class, and bit-identical f64 sums on every rank.
__global__ void k_reduce_gathered(const uint8_t* __restrict__ gathered, uint8_t* __restrict__ arena, size_t arena_bytes, int n_ranks,
                                  size_t b0, size_t e0, size_t b1, size_t e1, size_t b2, size_t e2, size_t b3, size_t e3) {
    const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = b0 + t0; i < e0; i += stride) {  // existence / Option flags: max(u8)
        uint8_t v = 0;
        for (int r = 0; r < n_ranks; r++) v |= gathered[(size_t)r * arena_bytes + i];
        arena[i] = v;
    }
    for (size_t i = b1 / 8 + t0; i < e1 / 8; i += stride) {  // counts, integer sums: wrapping sum(u64)
        uint64_t v = 0;
        for (int r = 0; r < n_ranks; r++) v += ((const uint64_t*)(gathered + (size_t)r * arena_bytes))[i];
        ((uint64_t*)arena)[i] = v;
    }
    for (size_t i = b2 / 8 + t0; i < e2 / 8; i += stride) {  // f64 sums, folded in rank order
        double v = -0.0;  // the identity of the f64 sum cells (dev.cuh F64_NEG_ZERO_BITS)
        for (int r = 0; r < n_ranks; r++) v = __dadd_rn(v, ((const double*)(gathered + (size_t)r * arena_bytes))[i]);
        ((double*)arena)[i] = v;
    }
    for (size_t i = b3 / 8 + t0; i < e3 / 8; i += stride) {  // min / max on order-preserving codes: max(u64)
        uint64_t v = 0;
        for (int r = 0; r < n_ranks; r++) {
            uint64_t x = ((const uint64_t*)(gathered + (size_t)r * arena_bytes))[i];
            v = x > v ? x : v;
        }
        ((uint64_t*)arena)[i] = v;
    }
}

#define AGREE_MAX (3 * TAGG_MAX_SCOPES + 8)  // words of a key-domain agreement

struct NcclState {
    void* lib = nullptr;
    ncclComm_t comm = nullptr;
    // key-domain agreement: its own stream, staging in pinned / device memory (one agreement in flight per context)
    cudaStream_t agree_st = nullptr;
    cudaEvent_t agree_ev = nullptr;
    uint64_t* agree_host = nullptr;
    uint64_t* agree_dev = nullptr;
    size_t agree_n = 0;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static int nccl_load(NcclState* s) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        s->lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (s->lib) break;
    }
    if (!s->lib) return tagg_fail(TAGG_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define LOAD(field, sym)                                                                    \
    *(void**)(&s->field) = dlsym(s->lib, sym);                                             \
    if (!s->field) return tagg_fail(TAGG_ERR_NCCL, "libnccl lacks %s", sym);
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(Reduce, "ncclReduce")
    LOAD(AllGather, "ncclAllGather")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    return 0;
}

#define NCCL_TRY(s, expr)                                                                               \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != ncclSuccess) return tagg_fail(TAGG_ERR_NCCL, "%s failed: %s", #expr, (s)->GetErrorString(_r)); \
    } while (0)

extern "C" {

int tagg_comm_unique_id(uint8_t out[TAGG_UNIQUE_ID_BYTES]) {
    if (!out) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    static_assert(sizeof(ncclUniqueId) <= TAGG_UNIQUE_ID_BYTES, "ncclUniqueId larger than TAGG_UNIQUE_ID_BYTES");
    NcclState s;
    int rc = nccl_load(&s);
    if (rc) return rc;
    ncclUniqueId id;
    NCCL_TRY(&s, s.GetUniqueId(&id));
    memset(out, 0, TAGG_UNIQUE_ID_BYTES);
    memcpy(out, &id, sizeof(id));
    return 0;
}

int tagg_comm_init(tagg_ctx* ctx, const uint8_t id_bytes[TAGG_UNIQUE_ID_BYTES], int rank, int n_ranks) {
    if (!ctx || !id_bytes || n_ranks < 1 || rank < 0 || rank >= n_ranks) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_comm_init: bad argument");
    if (ctx->nccl) return tagg_fail(TAGG_ERR_BAD_ARG, "communicator already initialised");
    CUDA_TRY(cudaSetDevice(ctx->device));
    auto* s = new NcclState();
    int rc = nccl_load(s);
    if (rc) { delete s; return rc; }
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclResult_t r = s->CommInitRank(&s->comm, n_ranks, id, rank);
    if (r != ncclSuccess) {
        rc = tagg_fail(TAGG_ERR_NCCL, "ncclCommInitRank failed: %s", s->GetErrorString(r));
        delete s;
        return rc;
    }
    if (cudaStreamCreateWithFlags(&s->agree_st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->agree_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaHostAlloc((void**)&s->agree_host, AGREE_MAX * 8, cudaHostAllocDefault) != cudaSuccess ||
        cudaMalloc((void**)&s->agree_dev, AGREE_MAX * 8) != cudaSuccess) {
        rc = tagg_fail(TAGG_ERR_CUDA, "communicator staging allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        s->CommDestroy(s->comm);
        delete s;
        return rc;
    }
    ctx->nccl = s;
    ctx->rank = rank;
    ctx->n_ranks = n_ranks;
    return 0;
}

int tagg_comm_destroy(tagg_ctx* ctx) {
    if (!ctx || !ctx->nccl) return 0;
    auto* s = (NcclState*)ctx->nccl;
    cudaSetDevice(ctx->device);
    if (s->agree_st) { cudaStreamSynchronize(s->agree_st); cudaStreamDestroy(s->agree_st); }
    if (s->agree_ev) cudaEventDestroy(s->agree_ev);
    if (s->agree_host) cudaFreeHost(s->agree_host);
    if (s->agree_dev) cudaFree(s->agree_dev);
    if (s->comm) s->CommDestroy(s->comm);
    delete s;
    ctx->nccl = nullptr;
    ctx->n_ranks = 1;
    ctx->rank = 0;
    return 0;
}

}  // extern "C"

// All ranks must lay their bucket tables out identically: agree on every scope's key domain.
// dom[3*s + 0] = smallest key (or ~0 if this rank saw none), [1] = ~largest key, [2] = dense_ok, last word: 0 if this rank
// needs the exact f64 MIN / MAX path.  One min all-reduce on the communicator's own stream; the call returns at once
// (the pass can start on the previous agreement meanwhile, exec.cu) and comm_agree_wait delivers the vector.
int comm_agree_begin(ExecState& es, const std::vector<uint64_t>& dom) {
    auto* s = (NcclState*)es.ctx->nccl;
    if (!s) return tagg_fail(TAGG_ERR_NCCL, "no communicator");
    if (dom.size() > AGREE_MAX) return tagg_fail(TAGG_ERR_BAD_PLAN, "too many bucket scopes for a collective call");
    s->agree_n = dom.size();
    if (dom.empty()) return 0;
    memcpy(s->agree_host, dom.data(), dom.size() * 8);
    CUDA_TRY(cudaMemcpyAsync(s->agree_dev, s->agree_host, dom.size() * 8, cudaMemcpyHostToDevice, s->agree_st));
    NCCL_TRY(s, s->AllReduce(s->agree_dev, s->agree_dev, dom.size(), ncclUint64, ncclMin, s->comm, s->agree_st));
    CUDA_TRY(cudaMemcpyAsync(s->agree_host, s->agree_dev, dom.size() * 8, cudaMemcpyDeviceToHost, s->agree_st));
    CUDA_TRY(cudaEventRecord(s->agree_ev, s->agree_st));
    return 0;
}
int comm_agree_wait(ExecState& es, std::vector<uint64_t>& agreed) {
    auto* s = (NcclState*)es.ctx->nccl;
    if (!s) return tagg_fail(TAGG_ERR_NCCL, "no communicator");
    agreed.assign(s->agree_n, 0);
    if (!s->agree_n) return 0;
    CUDA_TRY(cudaEventSynchronize(s->agree_ev));
    memcpy(agreed.data(), s->agree_host, s->agree_n * 8);
    // later collectives of this call run on the call's stream: order them behind the agreement explicitly
    CUDA_TRY(cudaStreamWaitEvent(es.st, s->agree_ev, 0));
    return 0;
}

// In-place merge of the accumulators (dense scopes only: identical layout on every rank).  root < 0: every rank ends up
// with the merged tables (small arenas: one all-gather + a class-aware reduction in rank order on every rank — one NCCL
// call, bit-identical f64 sums everywhere; large ones: one all-reduce per reduction class).  root >= 0: ncclReduce per
// reduction class into `root` only, one group (north_star (6): "bucket tables merged by an NCCL reduce over NVLink").
int comm_merge_arena(ExecState& es, int root) {
    auto* s = (NcclState*)es.ctx->nccl;
    if (!s) return tagg_fail(TAGG_ERR_NCCL, "no communicator");
    const ncclDataType_t dt[4] = {ncclUint8, ncclUint64, ncclFloat64, ncclUint64};
    const ncclRedOp_t op[4] = {ncclMax, ncclSum, ncclSum, ncclMax};
    const size_t esz[4] = {1, 8, 8, 8};
    if (root >= 0) {
        NCCL_TRY(s, s->GroupStart());
        for (int c = 0; c < 4; c++) {
            size_t bytes = es.cls_end[c] - es.cls_begin[c];
            if (!bytes) continue;
            NCCL_TRY(s, s->Reduce(es.arena + es.cls_begin[c], es.arena + es.cls_begin[c], bytes / esz[c], dt[c], op[c], root, s->comm, es.st));
        }
        NCCL_TRY(s, s->GroupEnd());
        return 0;
    }
    if (es.arena_bytes * (size_t)es.ctx->n_ranks <= (256u << 20)) {
        uint8_t* gathered = nullptr;
        CUDA_TRY(cudaMallocAsync((void**)&gathered, es.arena_bytes * (size_t)es.ctx->n_ranks, es.st));
        es.temps.push_back(gathered);
        NCCL_TRY(s, s->AllGather(es.arena, gathered, es.arena_bytes, ncclUint8, s->comm, es.st));
        size_t work = (es.cls_end[3] - es.cls_begin[0]) / 8 + 1;
        unsigned blocks = (unsigned)std::min<size_t>((work + 255) / 256, (size_t)es.ctx->sm_count * 8);
        k_reduce_gathered<<<blocks, 256, 0, es.st>>>(gathered, es.arena, es.arena_bytes, es.ctx->n_ranks, es.cls_begin[0], es.cls_end[0],
                                                     es.cls_begin[1], es.cls_end[1], es.cls_begin[2], es.cls_end[2], es.cls_begin[3], es.cls_end[3]);
        CUDA_TRY(cudaGetLastError());
        es.ctx->launches++;
        es.n_launches++;
        return 0;
    }
    // the arena is laid out by reduction class (exec.cu layout_arena): one all-reduce per class
    NCCL_TRY(s, s->GroupStart());
    for (int c = 0; c < 4; c++) {
        size_t bytes = es.cls_end[c] - es.cls_begin[c];
        if (!bytes) continue;
        NCCL_TRY(s, s->AllReduce(es.arena + es.cls_begin[c], es.arena + es.cls_begin[c], bytes / esz[c], dt[c], op[c], s->comm, es.st));
    }
    NCCL_TRY(s, s->GroupEnd());
    return 0;
}

// ---- exchange of compact results -----------------------------------------------------------------------------------
// Hashed bucket tables (slot positions differ from rank to rank) and percentile summaries cannot be reduced cell by cell:
// every rank compacts its own result, the compact results are all-gathered (sizes first, then the padded byte images)
// and folded on the host with PreparedAgg::merge in rank order (searcher.rs:93-96) — identical fruit on every rank.
namespace {
struct Writer {
    std::vector<uint8_t> b;
    void u64(uint64_t v) { size_t at = b.size(); b.resize(at + 8); memcpy(b.data() + at, &v, 8); }
    void raw(const void* p, size_t n) { size_t at = b.size(); b.resize(at + ((n + 7) & ~(size_t)7)); if (n) memcpy(b.data() + at, p, n); }
};
struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    uint64_t u64() { uint64_t v = 0; if (p + 8 > end) { ok = false; return 0; } memcpy(&v, p, 8); p += 8; return v; }
    template <typename T> void raw(std::vector<T>& out, size_t n) {
        const size_t bytes = n * sizeof(T), padded = (bytes + 7) & ~(size_t)7;
        if (p + padded > end) { ok = false; return; }
        out.resize(n);
        if (bytes) memcpy(out.data(), p, bytes);
        p += padded;
    }
};
void serialise(const tagg_result& r, Writer& w) {
    w.u64(r.scopes.size());
    for (auto& s : r.scopes) { w.u64(s.keys.size()); w.raw(s.keys.data(), s.keys.size() * 8); w.raw(s.parents.data(), s.parents.size() * 4); }
    w.u64(r.slots.size());
    for (auto& s : r.slots) { w.u64(s.values.size()); w.raw(s.values.data(), s.values.size() * 8); w.raw(s.seen.data(), s.seen.size()); }
    w.u64(r.pcts.size());
    for (auto& mp : r.pcts) {
        w.u64(mp.size());
        for (auto& kv : mp) {
            w.u64(kv.first); w.u64(kv.second.n_total); w.u64(kv.second.ranks.size());
            w.raw(kv.second.ranks.data(), kv.second.ranks.size() * 8);
            w.raw(kv.second.value_bits.data(), kv.second.value_bits.size() * 8);
        }
    }
}
bool deserialise(Reader& rd, tagg_result& r) {
    r.scopes.resize(rd.u64());
    for (auto& s : r.scopes) { size_t n = rd.u64(); rd.raw(s.keys, n); rd.raw(s.parents, n); if (!rd.ok) return false; }
    r.slots.resize(rd.u64());
    for (auto& s : r.slots) { size_t n = rd.u64(); rd.raw(s.values, n); rd.raw(s.seen, n); if (!rd.ok) return false; }
    r.pcts.resize(rd.u64());
    for (auto& mp : r.pcts) {
        size_t ne = rd.u64();
        for (size_t i = 0; i < ne && rd.ok; i++) {
            uint64_t bucket = rd.u64();
            PctSummary& ps = mp[bucket];
            ps.n_total = rd.u64();
            size_t np = rd.u64();
            rd.raw(ps.ranks, np);
            rd.raw(ps.value_bits, np);
        }
    }
    return rd.ok;
}
}  // namespace

int comm_merge_results(ExecState& es, tagg_result* res) {
    auto* s = (NcclState*)es.ctx->nccl;
    if (!s) return tagg_fail(TAGG_ERR_NCCL, "no communicator");
    const int nr = es.ctx->n_ranks;
    res->materialize();
    res->merged_elsewhere = 0;
    Writer w;
    serialise(*res, w);
    // sizes, then the padded images
    uint64_t* d_sizes = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&d_sizes, (size_t)nr * 8, es.st));
    es.temps.push_back(d_sizes);
    const uint64_t mine = w.b.size();
    CUDA_TRY(cudaMemcpyAsync(d_sizes + es.ctx->rank, &mine, 8, cudaMemcpyHostToDevice, es.st));
    NCCL_TRY(s, s->AllGather(d_sizes + es.ctx->rank, d_sizes, 1, ncclUint64, s->comm, es.st));
    std::vector<uint64_t> sizes(nr);
    CUDA_TRY(cudaMemcpyAsync(sizes.data(), d_sizes, (size_t)nr * 8, cudaMemcpyDeviceToHost, es.st));
    CUDA_TRY(cudaStreamSynchronize(es.st));
    size_t slot = 0;
    for (uint64_t x : sizes) slot = std::max<size_t>(slot, x);
    slot = (slot + 255) & ~(size_t)255;
    if (!slot) return 0;
    uint8_t* d_all = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&d_all, slot * (size_t)nr, es.st));
    es.temps.push_back(d_all);
    CUDA_TRY(cudaMemcpyAsync(d_all + slot * (size_t)es.ctx->rank, w.b.data(), w.b.size(), cudaMemcpyHostToDevice, es.st));
    NCCL_TRY(s, s->AllGather(d_all + slot * (size_t)es.ctx->rank, d_all, slot, ncclUint8, s->comm, es.st));
    std::vector<uint8_t> all(slot * (size_t)nr);
    CUDA_TRY(cudaMemcpyAsync(all.data(), d_all, all.size(), cudaMemcpyDeviceToHost, es.st));
    CUDA_TRY(cudaStreamSynchronize(es.st));
    tagg_result merged;
    merged.meta = res->meta;
    for (int r = 0; r < nr; r++) {
        tagg_result part;
        part.meta = res->meta;
        Reader rd{all.data() + slot * (size_t)r, all.data() + slot * (size_t)r + sizes[r]};
        if (!deserialise(rd, part)) return tagg_fail(TAGG_ERR_NCCL, "malformed result image from rank %d", r);
        if (r == 0) { merged.scopes = std::move(part.scopes); merged.slots = std::move(part.slots); merged.pcts = std::move(part.pcts); }
        else {
            int rc = result_merge(&merged, &part);
            if (rc) return rc;
        }
    }
    res->scopes = std::move(merged.scopes);
    res->slots = std::move(merged.slots);
    res->pcts = std::move(merged.pcts);
    return 0;
}

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

#include "exec.h"

struct NcclState {
    void* lib = nullptr;
    ncclComm_t comm = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
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
    ctx->nccl = s;
    ctx->rank = rank;
    ctx->n_ranks = n_ranks;
    return 0;
}

int tagg_comm_destroy(tagg_ctx* ctx) {
    if (!ctx || !ctx->nccl) return 0;
    auto* s = (NcclState*)ctx->nccl;
    if (s->comm) s->CommDestroy(s->comm);
    delete s;
    ctx->nccl = nullptr;
    ctx->n_ranks = 1;
    ctx->rank = 0;
    return 0;
}

}  // extern "C"

// All ranks must lay their bucket tables out identically: agree on every scope's key domain.
// in/out: dom[3*s + 0] = smallest key (or ~0 if this rank saw none), [1] = ~largest key, [2] = dense_ok.
int comm_agree_domains(ExecState& es, std::vector<uint64_t>& dom) {
    auto* s = (NcclState*)es.ctx->nccl;
    if (!s || dom.empty()) return 0;
    uint64_t* d = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&d, dom.size() * 8, es.st));
    CUDA_TRY(cudaMemcpyAsync(d, dom.data(), dom.size() * 8, cudaMemcpyHostToDevice, es.st));
    NCCL_TRY(s, s->AllReduce(d, d, dom.size(), ncclUint64, ncclMin, s->comm, es.st));
    CUDA_TRY(cudaMemcpyAsync(dom.data(), d, dom.size() * 8, cudaMemcpyDeviceToHost, es.st));
    CUDA_TRY(cudaStreamSynchronize(es.st));
    cudaFreeAsync(d, es.st);
    return 0;
}

int comm_merge_arena(ExecState& es) {
    auto* s = (NcclState*)es.ctx->nccl;
    if (!s) return tagg_fail(TAGG_ERR_NCCL, "no communicator");
    const PlanMeta& m = *es.meta;
    for (auto& L : es.scopes)
        if (L.mode != SCOPE_DENSE)
            return tagg_fail(TAGG_ERR_UNSUPPORTED, "collective merge of hashed bucket tables is not implemented yet (dense key domains only)");
    if (!m.pct_node.empty()) return tagg_fail(TAGG_ERR_UNSUPPORTED, "collective merge of percentiles is not implemented yet");
    NCCL_TRY(s, s->GroupStart());
    for (size_t sc = 1; sc < es.scopes.size(); sc++) {
        const ScopeLayout& L = es.scopes[sc];
        NCCL_TRY(s, s->AllReduce(es.arena + L.off_present, es.arena + L.off_present, L.capacity, ncclUint8, ncclMax, s->comm, es.st));
    }
    for (size_t k = 0; k < es.slots.size(); k++) {
        const SlotLayout& SL = es.slots[k];
        const tagg_node& nd = m.nodes[m.slot_node[k]];
        void* acc = es.arena + SL.off_acc;
        if (nd.op == TAGG_OP_COUNT || (nd.op == TAGG_OP_SUM && nd.kind != TAGG_F64))
            NCCL_TRY(s, s->AllReduce(acc, acc, SL.capacity, ncclUint64, ncclSum, s->comm, es.st));
        else if (nd.op == TAGG_OP_SUM)
            NCCL_TRY(s, s->AllReduce(acc, acc, SL.capacity, ncclFloat64, ncclSum, s->comm, es.st));
        else
            NCCL_TRY(s, s->AllReduce(acc, acc, SL.capacity, ncclUint64, ncclMax, s->comm, es.st));
        NCCL_TRY(s, s->AllReduce(es.arena + SL.off_seen, es.arena + SL.off_seen, SL.capacity, ncclUint8, ncclMax, s->comm, es.st));
    }
    NCCL_TRY(s, s->GroupEnd());
    CUDA_TRY(cudaStreamSynchronize(es.st));
    return 0;
}

// api.cu — C ABI entry points: library, context, plans and the execute orchestration.
//
// `tagg_execute` is the replacement of the reference's `collect_segment`
// (src/searcher.rs:27-51) for a batch of segments folded into one fruit
// (Executor::SingleThread, src/searcher.rs:66-78).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <set>

#include "exec.h"
#include "host.h"

// ---- errors --------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int tagg_fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

cudaStream_t tagg_ctx::acquire_stream() {
    std::lock_guard<std::mutex> g(mu);
    if (!stream_pool.empty()) {
        cudaStream_t s = stream_pool.back();
        stream_pool.pop_back();
        return s;
    }
    cudaStream_t s = nullptr;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    return s;
}
void tagg_ctx::release_stream(cudaStream_t s) {
    std::lock_guard<std::mutex> g(mu);
    stream_pool.push_back(s);
}

CallRes* tagg_ctx::acquire_call() {
    {
        std::lock_guard<std::mutex> g(mu);
        if (!call_pool.empty()) {
            CallRes* c = call_pool.back();
            call_pool.pop_back();
            c->pinned_used = 0;
            return c;
        }
    }
    auto* c = new CallRes();
    c->st = acquire_stream();
    c->st2 = acquire_stream();
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    for (auto& e : c->chunk_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->join_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->chain_ev, cudaEventDisableTiming);
    c->pinned_bytes = 8 << 20;  // arenas up to this size come back whole in one pinned copy
    if (cudaHostAlloc((void**)&c->pinned, c->pinned_bytes, cudaHostAllocDefault) != cudaSuccess) {
        c->pinned = nullptr;
        c->pinned_bytes = 0;
        cudaGetLastError();
    }
    return c;
}
void tagg_ctx::release_call(CallRes* c) {
    std::lock_guard<std::mutex> g(mu);
    call_pool.push_back(c);
}

// ---- plan analysis ---------------------------------------------------------------------------------
static int analyse(PlanMeta& m, uint32_t& pos, int parent, int scope, int depth) {
    uint32_t n = (uint32_t)m.nodes.size();
    if (pos >= n) return tagg_fail(TAGG_ERR_BAD_PLAN, "plan ends inside a sub-tree");
    if (depth > TAGG_MAX_DEPTH) return tagg_fail(TAGG_ERR_BAD_PLAN, "plan nests deeper than %d", TAGG_MAX_DEPTH);
    uint32_t me = pos++;
    const tagg_node& d = m.nodes[me];
    m.parent_node[me] = parent;
    m.scope_of[me] = scope;
    uint32_t want_lo = 0, want_hi = 0;
    bool reads_col = true;
    switch (d.op) {
        case TAGG_OP_TUPLE: want_lo = 2; want_hi = 10; reads_col = false; break;  // tuple.rs:73-81
        case TAGG_OP_COUNT: reads_col = false; break;
        case TAGG_OP_SUM:
            if (d.kind != TAGG_U64 && d.kind != TAGG_I64 && d.kind != TAGG_F64)
                return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: sum_agg exists for u64/i64/f64 only (sum.rs:146-158)", me);
            break;
        case TAGG_OP_MIN:
        case TAGG_OP_MAX:
            if (d.kind > TAGG_DATE) return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: bad kind", me);
            break;
        case TAGG_OP_PERCENTILES:
            if (d.kind != TAGG_F64) return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: percentiles_agg exists for f64 only (percentile.rs:130-138)", me);
            break;
        case TAGG_OP_TERMS:
            if (d.kind != TAGG_U64 && d.kind != TAGG_I64)
                return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: terms_agg exists for u64/i64 only (terms.rs:185-195)", me);
            want_lo = want_hi = 1;
            break;
        case TAGG_OP_HISTOGRAM:
            // f64 keys: histogram_agg_f64 (histogram.rs:9-21); i64 / date keys: date_histogram (README.md:41, beyond the reference)
            if ((d.kind != TAGG_F64 && d.kind != TAGG_I64 && d.kind != TAGG_DATE) || d.multi)
                return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: histogram_agg reads a single-valued f64 (or i64 / date) field (histogram.rs:9-21)", me);
            want_lo = want_hi = 1;
            break;
        case TAGG_OP_FILTER:
            if (d.aux >= TAGG_MAX_FILTERS) return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: at most %d filter_agg nodes per plan", me, TAGG_MAX_FILTERS);
            m.n_filters = std::max(m.n_filters, d.aux + 1);
            want_lo = want_hi = 1;
            reads_col = false;
            break;
        case TAGG_OP_POST_FILTER:
            if (d.kind > TAGG_DATE) return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: bad kind", me);
            if (d.pred != TAGG_PRED_RANGE && d.pred != TAGG_PRED_LUT)
                return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: post_filter needs a RANGE or LUT predicate", me);
            if (d.pred == TAGG_PRED_LUT) {
                if (d.aux >= m.blobs.size()) return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: LUT blob %u missing", me, d.aux);
                if (m.blobs[d.aux].size() * 8 < d.u1) return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: LUT blob shorter than u1 bits", me);
            }
            want_lo = want_hi = 1;
            break;
        default: return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u: unknown op %u", me, d.op);
    }
    if (d.n_children < want_lo || d.n_children > want_hi)
        return tagg_fail(TAGG_ERR_BAD_PLAN, "node %u (op %u): %u children, expected %u..%u", me, d.op, d.n_children, want_lo, want_hi);
    if (reads_col) {
        int slot = -1;
        int at = 0;
        for (auto& r : m.colrefs) {
            if (r.field_id == d.field_id && r.multi == (d.multi ? 1 : 0)) { slot = at; break; }
            at += r.multi ? 2 : 1;
        }
        if (slot < 0) {
            slot = at;
            if (at + (d.multi ? 2 : 1) > TAGG_MAX_COLS) return tagg_fail(TAGG_ERR_BAD_PLAN, "plan reads more than %d device columns", TAGG_MAX_COLS);
            m.colrefs.push_back({d.field_id, d.multi ? 1 : 0});
        }
        m.col_slot[me] = slot;
    }
    int child_scope = scope;
    if (d.op == TAGG_OP_TERMS || d.op == TAGG_OP_HISTOGRAM) {
        if (m.scope_node.size() >= TAGG_MAX_SCOPES) return tagg_fail(TAGG_ERR_BAD_PLAN, "more than %d bucket scopes", TAGG_MAX_SCOPES);
        child_scope = (int)m.scope_node.size();
        m.own_scope[me] = child_scope;
        m.scope_node.push_back((int)me);
        m.scope_parent.push_back(scope);
    } else if (d.op >= TAGG_OP_COUNT && d.op <= TAGG_OP_MAX) {
        m.slot_of[me] = (int)m.slot_node.size();
        m.slot_node.push_back((int)me);
    } else if (d.op == TAGG_OP_PERCENTILES) {
        if (m.pct_node.size() >= 4) return tagg_fail(TAGG_ERR_UNSUPPORTED, "more than 4 percentiles aggregations in a plan");
        m.pct_of[me] = (int)m.pct_node.size();
        m.pct_node.push_back((int)me);
    }
    for (uint32_t i = 0; i < d.n_children; i++) {
        int rc = analyse(m, pos, (int)me, child_scope, depth + 1);
        if (rc) return rc;
    }
    m.end[me] = (uint16_t)pos;
    return 0;
}

// contexts that are alive (results are recycled into their context's pool on free)
static std::mutex g_live_mu;
static std::set<tagg_ctx*> g_live_ctx;

extern "C" {

uint32_t tagg_abi_version(void) { return TAGG_ABI_VERSION; }
const char* tagg_last_error(void) { return g_err; }

int tagg_device_count(int* out) {
    if (!out) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *out = 0;
        return tagg_fail(TAGG_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *out = n;
    return 0;
}

int tagg_ctx_create(int device, tagg_ctx** out) {
    if (!out) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return tagg_fail(TAGG_ERR_NO_DEVICE, "no CUDA device (%s); the aggregation hot path has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return tagg_fail(TAGG_ERR_BAD_ARG, "device %d out of range (0..%d)", device, n - 1);
    CUDA_TRY(cudaSetDevice(device));
    auto* c = new tagg_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    {   // keep freed stream-ordered allocations cached in the pool: by default they go back to the driver at
        // every synchronisation and the next query pays for mapping them again
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if (prop.major < 10) {
        delete c;
        return tagg_fail(TAGG_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    }
    {
        std::lock_guard<std::mutex> live(g_live_mu);
        g_live_ctx.insert(c);
    }
    *out = c;
    return 0;
}

int tagg_ctx_destroy(tagg_ctx* ctx) {
    if (!ctx) return 0;
    {
        std::lock_guard<std::mutex> live(g_live_mu);
        g_live_ctx.erase(ctx);
    }
    for (auto r : ctx->result_pool) delete r;
    ctx->result_pool.clear();
    cudaSetDevice(ctx->device);
    tagg_comm_destroy(ctx);
    for (auto c : ctx->call_pool) {
        for (auto& b : c->blocks) cudaFree(b.p);
        cudaStreamDestroy(c->st);
        cudaStreamDestroy(c->st2);
        for (auto e : c->chunk_ev) cudaEventDestroy(e);
        cudaEventDestroy(c->join_ev);
        cudaEventDestroy(c->chain_ev);
        cudaEventDestroy(c->ev0);
        cudaEventDestroy(c->ev1);
        if (c->pinned) cudaFreeHost(c->pinned);
        delete c;
    }
    if (ctx->pct_sched_dev) cudaFree(ctx->pct_sched_dev);
    for (auto s : ctx->stream_pool) cudaStreamDestroy(s);
    if (ctx->timer0) { cudaEventDestroy(ctx->timer0); cudaEventDestroy(ctx->timer1); }
    delete ctx;
    return 0;
}

int tagg_ctx_device(const tagg_ctx* ctx, int* out) {
    if (!ctx || !out) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    *out = ctx->device;
    return 0;
}

int tagg_ctx_synchronize(tagg_ctx* ctx) {
    if (!ctx) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

int tagg_ctx_set_path(tagg_ctx* ctx, int path) {
    if (!ctx || path < 0 || path > 2) return tagg_fail(TAGG_ERR_BAD_ARG, "bad path");
    ctx->path = path;
    return 0;
}

int tagg_ctx_timer_start(tagg_ctx* ctx) {
    if (!ctx) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (!ctx->timer0) { CUDA_TRY(cudaEventCreate(&ctx->timer0)); CUDA_TRY(cudaEventCreate(&ctx->timer1)); }
    CUDA_TRY(cudaDeviceSynchronize());
    CallRes* cr = ctx->acquire_call();  // LIFO pool: the same stream the next execute will use
    cudaError_t e = cudaEventRecord(ctx->timer0, cr->st);
    ctx->release_call(cr);
    if (e != cudaSuccess) return tagg_fail(TAGG_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(e));
    return 0;
}

int tagg_ctx_timer_stop(tagg_ctx* ctx, double* ms) {
    if (!ctx || !ms || !ctx->timer0) return tagg_fail(TAGG_ERR_BAD_ARG, "timer not started");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CallRes* cr = ctx->acquire_call();
    cudaError_t e = cudaEventRecord(ctx->timer1, cr->st);
    ctx->release_call(cr);
    if (e != cudaSuccess) return tagg_fail(TAGG_ERR_CUDA, "cudaEventRecord failed: %s", cudaGetErrorString(e));
    CUDA_TRY(cudaEventSynchronize(ctx->timer1));
    float f = 0;
    CUDA_TRY(cudaEventElapsedTime(&f, ctx->timer0, ctx->timer1));
    *ms = f;
    return 0;
}

int tagg_ctx_launch_count(const tagg_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    *out = ctx->launches.load();
    return 0;
}

int tagg_plan_create(tagg_ctx* ctx, const tagg_node* nodes, uint32_t n_nodes, const tagg_blob* blobs, uint32_t n_blobs,
                     tagg_plan** out) {
    if (!ctx || !nodes || !out || n_nodes == 0) return tagg_fail(TAGG_ERR_BAD_ARG, "tagg_plan_create: null argument");
    if (n_nodes > TAGG_MAX_NODES) return tagg_fail(TAGG_ERR_BAD_PLAN, "plan has %u nodes, at most %d supported", n_nodes, TAGG_MAX_NODES);
    auto m = std::make_shared<PlanMeta>();
    m->nodes.assign(nodes, nodes + n_nodes);
    for (uint32_t i = 0; i < n_blobs; i++) {
        if (!blobs || !blobs[i].data) return tagg_fail(TAGG_ERR_BAD_ARG, "null blob");
        m->blobs.emplace_back(blobs[i].data, blobs[i].data + blobs[i].len);
    }
    m->end.assign(n_nodes, 0);
    m->parent_node.assign(n_nodes, -1);
    m->scope_of.assign(n_nodes, 0);
    m->own_scope.assign(n_nodes, -1);
    m->slot_of.assign(n_nodes, -1);
    m->pct_of.assign(n_nodes, -1);
    m->col_slot.assign(n_nodes, -1);
    m->scope_node.push_back(-1);  // root scope
    m->scope_parent.push_back(-1);
    uint32_t pos = 0;
    int rc = analyse(*m, pos, -1, 0, 0);
    if (rc) return rc;
    if (pos != n_nodes) return tagg_fail(TAGG_ERR_BAD_PLAN, "%u trailing nodes after the root sub-tree", n_nodes - pos);
    CUDA_TRY(cudaSetDevice(ctx->device));
    auto* p = new tagg_plan();
    p->ctx = ctx;
    p->meta = m;
    for (auto& b : m->blobs) {
        uint8_t* d = nullptr;
        cudaError_t e = cudaMalloc(&d, b.size() + 16);
        if (e == cudaSuccess) e = cudaMemcpy(d, b.data(), b.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            tagg_plan_destroy(p);
            return tagg_fail(TAGG_ERR_CUDA, "LUT upload failed: %s", cudaGetErrorString(e));
        }
        p->d_blobs.push_back(d);
    }
    *out = p;
    return 0;
}

int tagg_plan_set_readout(tagg_plan* plan, int readout) {
    if (!plan || (readout != TAGG_READOUT_EAGER && readout != TAGG_READOUT_LAZY)) return tagg_fail(TAGG_ERR_BAD_ARG, "bad readout mode");
    plan->readout = readout;
    return 0;
}

int tagg_plan_destroy(tagg_plan* plan) {
    if (!plan) return 0;
    cudaSetDevice(plan->ctx->device);
    for (auto d : plan->d_blobs) cudaFree(d);
    delete plan;
    return 0;
}

int tagg_execute(const tagg_plan* plan, const tagg_segment_input* inputs, uint32_t n_inputs, tagg_result** out) {
    return exec_run(plan, inputs, n_inputs, 0, -1, out);
}

int tagg_execute_begin(const tagg_plan* plan, const tagg_segment_input* inputs, uint32_t n_inputs, tagg_pending** out) {
    return exec_begin(plan, inputs, n_inputs, out);
}

int tagg_pending_wait(tagg_pending* pending, tagg_result** out) { return exec_wait(pending, out); }

int tagg_execute_collective(const tagg_plan* plan, const tagg_segment_input* inputs, uint32_t n_inputs, tagg_result** out) {
    return exec_run(plan, inputs, n_inputs, 1, -1, out);
}

int tagg_execute_reduce(const tagg_plan* plan, const tagg_segment_input* inputs, uint32_t n_inputs, int root, tagg_result** out) {
    return exec_run(plan, inputs, n_inputs, 2, root, out);
}

int tagg_result_is_local(const tagg_result* res, int* out) {
    if (!res || !out) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    *out = res->merged_elsewhere ? 0 : 1;
    return 0;
}

int tagg_result_free(tagg_result* res) {
    if (res && res->ctx) {  // recycle (bounded): the arrays keep their capacity for the next result
        std::lock_guard<std::mutex> live(g_live_mu);
        tagg_ctx* ctx = res->ctx;
        if (g_live_ctx.count(ctx)) {  // (a result may outlive its context)
            res->release_device();
            res->meta.reset();
            res->pcts.clear();
            std::lock_guard<std::mutex> g(ctx->mu);
            if (ctx->result_pool.size() < 4) { ctx->result_pool.push_back(res); return 0; }
        }
    }
    delete res;
    return 0;
}

int tagg_result_merge(tagg_result* dst, const tagg_result* src) {
    if (!dst || !src) return tagg_fail(TAGG_ERR_BAD_ARG, "null argument");
    return result_merge(dst, src);
}

}  // extern "C"

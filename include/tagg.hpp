// tagg.hpp — typed C++17 host façade over the C ABI (include/tagg.h), mirroring the reference's
// Rust API for the hot path: the trait triple of src/agg.rs:10-36 becomes a statically typed
// aggregation tree whose `Fruit` type is computed at compile time, exactly like the reference's
// generics + tuples (src/tuple.rs:5-81):
//
//   auto agg  = filter_agg(term_query_u64(status, 0),
//                          tuple(count_agg(), terms_agg_u64(category, tuple(count_agg(), min_agg_f64(price)))));
//   auto fruit = searcher.agg_search(all_query(), agg);          // std::tuple<uint64_t, Terms<uint64_t, std::tuple<uint64_t, std::optional<double>>>>
//
// The reference is Rust; no Rust toolchain exists in the build image, so the compiled host side above
// the C ABI is C++ (the Rust binding a maintainer would add is in INTEGRATION.md).  Everything here is
// host glue: lowering to the flat plan (Agg::emit), docset hand-over, fruit decoding (Agg::read).
// All computation happens in libtagg.so's CUDA kernels; errors surface as tagg::Error (no fallback).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include "tagg.h"

namespace tagg {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error("tagg status " + std::to_string(s) + ": " + m), status(s) {}
};
// tantivy's FastFieldNotAvailableError (reference sum.rs:50-55, terms.rs:76-81, ...)
struct FastFieldNotAvailableError : Error { using Error::Error; };
inline void check(int status) {
    if (status == TAGG_OK) return;
    if (status == TAGG_ERR_NO_SUCH_COLUMN) throw FastFieldNotAvailableError(status, tagg_last_error());
    throw Error(status, tagg_last_error());
}

using Field = uint32_t;  // tantivy::schema::Field

// ---- value <-> code (tantivy common::{f64_to_u64, i64_to_u64}) -----------------------------------
inline uint64_t f64_to_code(double v) { uint64_t b; std::memcpy(&b, &v, 8); return (b >> 63) == 0 ? b ^ (1ull << 63) : ~b; }
inline uint64_t i64_to_code(int64_t v) { return (uint64_t)v ^ (1ull << 63); }
template <class T> struct Kind;
template <> struct Kind<uint64_t> { static constexpr int id = TAGG_U64; static uint64_t code(uint64_t v) { return v; } static uint64_t from_bits(uint64_t b) { return b; } };
template <> struct Kind<int64_t> { static constexpr int id = TAGG_I64; static uint64_t code(int64_t v) { return i64_to_code(v); } static int64_t from_bits(uint64_t b) { return (int64_t)b; } };
template <> struct Kind<double> { static constexpr int id = TAGG_F64; static uint64_t code(double v) { return f64_to_code(v); } static double from_bits(uint64_t b) { double d; std::memcpy(&d, &b, 8); return d; } };
struct DateTime { int64_t timestamp; bool operator==(const DateTime& o) const { return timestamp == o.timestamp; } };  // chrono::DateTime<Utc>, second resolution
template <> struct Kind<DateTime> { static constexpr int id = TAGG_DATE; static uint64_t code(DateTime v) { return i64_to_code(v.timestamp); } static DateTime from_bits(uint64_t b) { return DateTime{(int64_t)b}; } };

// ---- handles ----------------------------------------------------------------------------------------
class Context {
  public:
    explicit Context(int device = 0) { check(tagg_ctx_create(device, &h_)); }
    ~Context() { tagg_ctx_destroy(h_); }
    Context(const Context&) = delete;
    tagg_ctx* get() const { return h_; }
    void set_path(int p) { check(tagg_ctx_set_path(h_, p)); }
  private:
    tagg_ctx* h_ = nullptr;
};

// A segment's fast fields resident in HBM (stand-in for tantivy::SegmentReader on this path).
class Segment {
  public:
    Segment(const Context& ctx, uint32_t max_doc) : max_doc_(max_doc) { check(tagg_segment_create(ctx.get(), max_doc, &h_)); }
    ~Segment() { tagg_segment_destroy(h_); }
    Segment(const Segment&) = delete;
    template <class T> void add_column(Field f, const std::vector<T>& values) {
        std::vector<uint64_t> codes(values.size());
        for (size_t i = 0; i < values.size(); i++) codes[i] = Kind<T>::code(values[i]);
        check(tagg_column_upload_codes(h_, f, Kind<T>::id, codes.data(), codes.size()));
    }
    void add_column_bytes(Field f, int kind, const uint8_t* bytes, size_t len) { check(tagg_column_upload(h_, f, kind, bytes, len)); }
    template <class T> void add_multicolumn(Field f, const std::vector<std::vector<T>>& lists) {
        std::vector<uint64_t> off(lists.size() + 1, 0), codes;
        for (size_t i = 0; i < lists.size(); i++) {
            for (auto& v : lists[i]) codes.push_back(Kind<T>::code(v));
            off[i + 1] = codes.size();
        }
        check(tagg_multicolumn_upload_codes(h_, f, Kind<T>::id, off.data(), off.size(), codes.data(), codes.size()));
    }
    void set_deletes(const std::vector<uint8_t>& bitset) { check(tagg_segment_set_deletes(h_, bitset.data(), bitset.size())); }
    tagg_segment* get() const { return h_; }
    uint32_t max_doc() const { return max_doc_; }
  private:
    tagg_segment* h_ = nullptr;
    uint32_t max_doc_;
};

// ---- queries: what a tantivy Weight::scorer(segment) yields, as a docset ---------------------------------
struct Query {
    virtual ~Query() = default;
    virtual tagg_docset docset(const Segment& seg, size_t ord) const = 0;
};
struct AllQuery : Query {
    tagg_docset docset(const Segment&, size_t) const override { tagg_docset d{}; d.kind = TAGG_DOCSET_ALL; return d; }
};
inline AllQuery all_query() { return AllQuery(); }
// TermQuery / RangeQuery on an INDEXED|FAST field, evaluated on the device from the fast field
struct ColumnRangeQuery : Query {
    Field field; uint64_t lo, hi;
    ColumnRangeQuery(Field f, uint64_t l, uint64_t h) : field(f), lo(l), hi(h) {}
    tagg_docset docset(const Segment&, size_t) const override {
        tagg_docset d{}; d.kind = TAGG_DOCSET_COLUMN_RANGE; d.field_id = field; d.lo = lo; d.hi = hi; return d;
    }
};
inline ColumnRangeQuery term_query_u64(Field f, uint64_t v) { return ColumnRangeQuery(f, v, v); }
inline ColumnRangeQuery range_query_f64(Field f, double lo, double hi_excl) { return ColumnRangeQuery(f, f64_to_code(lo), f64_to_code(hi_excl) - 1); }  // RangeQuery::new_f64(f, lo..hi)
struct BitsetQuery : Query {  // a scorer drained into per-segment bitsets (tantivy BitSetDocSet)
    std::vector<std::vector<uint8_t>> per_segment;
    tagg_docset docset(const Segment&, size_t ord) const override {
        tagg_docset d{}; d.kind = TAGG_DOCSET_BITSET; d.data = per_segment[ord].data(); d.n = per_segment[ord].size(); return d;
    }
};
struct DocIdsQuery : Query {  // a scorer drained into sorted doc-id lists (TermScorer postings)
    std::vector<std::vector<uint32_t>> per_segment;
    tagg_docset docset(const Segment&, size_t ord) const override {
        tagg_docset d{}; d.kind = TAGG_DOCSET_SORTED_IDS; d.data = per_segment[ord].data(); d.n = per_segment[ord].size(); return d;
    }
};

// ---- plan builder / result reader -------------------------------------------------------------------------
struct PlanBuilder {
    std::vector<tagg_node> nodes;
    std::vector<const Query*> filters;
    uint32_t emit(tagg_node n) { nodes.push_back(n); return (uint32_t)nodes.size() - 1; }
};

class ResultReader {
  public:
    explicit ResultReader(tagg_result* r) : r_(r) {}
    ~ResultReader() { tagg_result_free(r_); }
    ResultReader(const ResultReader&) = delete;
    struct Scope { std::vector<uint64_t> keys; std::vector<uint32_t> parents; };
    struct Metric { std::vector<uint64_t> values; std::vector<uint8_t> seen; };
    const Scope& scope(uint32_t node) const {
        auto it = scopes_.find(node);
        if (it != scopes_.end()) return it->second;
        uint64_t n = 0;
        check(tagg_result_scope_len(r_, node, &n));
        Scope s; s.keys.resize(n); s.parents.resize(n);
        check(tagg_result_scope_read(r_, node, s.keys.data(), s.parents.data(), n));
        return scopes_.emplace(node, std::move(s)).first->second;
    }
    const Metric& metric(uint32_t node) const {
        auto it = metrics_.find(node);
        if (it != metrics_.end()) return it->second;
        uint64_t n = 0;
        check(tagg_result_metric_len(r_, node, &n));
        Metric m; m.values.resize(n); m.seen.resize(n);
        check(tagg_result_metric_read(r_, node, m.values.data(), m.seen.data(), n));
        return metrics_.emplace(node, std::move(m)).first->second;
    }
    tagg_result* get() const { return r_; }
  private:
    tagg_result* r_;
    mutable std::unordered_map<uint32_t, Scope> scopes_;
    mutable std::unordered_map<uint32_t, Metric> metrics_;
};

// ---- fruits ---------------------------------------------------------------------------------------------------
// Terms<K, T> — src/bucket/terms.rs:403-458
template <class K, class T>
class Terms {
  public:
    std::unordered_map<K, T> res;
    const T* get(const K& key) const { auto it = res.find(key); return it == res.end() ? nullptr : &it->second; }
    size_t len() const { return res.size(); }
    // top_k(k, sort_by): the k buckets with the largest sort key, descending; equal sort keys ascending by key
    template <class F> std::vector<std::pair<K, const T*>> top_k(size_t k, F sort_by) const {
        std::vector<std::pair<K, const T*>> v;
        for (auto& kv : res) v.emplace_back(kv.first, &kv.second);
        std::sort(v.begin(), v.end(), [&](auto& a, auto& b) {
            auto sa = sort_by(*a.second); auto sb = sort_by(*b.second);
            if (sb < sa) return true;
            if (sa < sb) return false;
            return a.first < b.first;
        });
        if (v.size() > k) v.resize(k);
        return v;
    }
};
// Histogram<T> — src/bucket/histogram.rs:156-181
template <class T>
class Histogram {
  public:
    double start = 0, interval = 0;
    std::map<uint64_t, T> bucket_map;  // BTreeMap<u64, T>
    std::vector<std::pair<double, const T*>> buckets() const {  // gap buckets materialised as nullptr (None)
        std::vector<std::pair<double, const T*>> out;
        bool first = true; uint64_t last = 0;
        for (auto& kv : bucket_map) {
            if (!first && kv.first - last > 1)
                for (uint64_t i = 0; i < kv.first - last - 1; i++) out.emplace_back((double)(last + i + 1) * interval + start, nullptr);
            out.emplace_back((double)kv.first * interval + start, &kv.second);
            last = kv.first; first = false;
        }
        return out;
    }
};
// Percentiles<f64> — src/metric/percentile.rs:152-177: exact order statistics (rank, value); percentile(q)
// answers with the stored statistic nearest the rank the reference's CKMS(0.01) targets.
class Percentiles {
  public:
    uint64_t n = 0;
    std::vector<uint64_t> ranks;
    std::vector<double> values;
    std::optional<double> percentile(double q) const {
        if (n == 0 || ranks.empty()) return std::nullopt;
        double nphi = q * (double)n;
        double inv = std::max(1.0, std::floor(2.0 * 0.01 * nphi));
        double kf = std::floor(nphi + inv / 2.0);
        uint64_t k = kf < 1.0 ? 1 : (kf > (double)n ? n : (uint64_t)kf);
        size_t i = std::lower_bound(ranks.begin(), ranks.end(), k) - ranks.begin();
        if (i == ranks.size()) i--;
        else if (i > 0 && ranks[i] != k && (k - ranks[i - 1]) <= (ranks[i] - k)) i--;
        return values[i];
    }
};

// ---- aggregation nodes: each has `Fruit`, `emit(PlanBuilder&)`, `read(reader, bucket)` --------------------------
struct CountAgg {  // count_agg() — src/metric/count.rs:7-9
    using Fruit = uint64_t;
    mutable uint32_t node = 0;
    void emit(PlanBuilder& pb) const { tagg_node n{}; n.op = TAGG_OP_COUNT; node = pb.emit(n); }
    Fruit read(const ResultReader& r, uint32_t bucket) const { return r.metric(node).values[bucket]; }
};
inline CountAgg count_agg() { return CountAgg(); }

template <class T, int OP, bool MULTI>
struct FoldAgg {  // sum / min / max: Fruit = Option<T> — src/metric/sum.rs, src/metric/minmax.rs
    using Fruit = std::optional<T>;
    Field field;
    mutable uint32_t node = 0;
    void emit(PlanBuilder& pb) const {
        tagg_node n{}; n.op = OP; n.kind = Kind<T>::id; n.multi = MULTI; n.field_id = field; node = pb.emit(n);
    }
    Fruit read(const ResultReader& r, uint32_t bucket) const {
        auto& m = r.metric(node);
        if (!m.seen[bucket]) return std::nullopt;
        return Kind<T>::from_bits(m.values[bucket]);
    }
};
#define TAGG_FOLD_CTOR(name, T, OP) \
    inline FoldAgg<T, OP, false> name(Field f) { return {f}; } \
    inline FoldAgg<T, OP, true> name##s(Field f) { return {f}; }
TAGG_FOLD_CTOR(sum_agg_u64, uint64_t, TAGG_OP_SUM) TAGG_FOLD_CTOR(sum_agg_i64, int64_t, TAGG_OP_SUM) TAGG_FOLD_CTOR(sum_agg_f64, double, TAGG_OP_SUM)
TAGG_FOLD_CTOR(min_agg_u64, uint64_t, TAGG_OP_MIN) TAGG_FOLD_CTOR(min_agg_i64, int64_t, TAGG_OP_MIN) TAGG_FOLD_CTOR(min_agg_f64, double, TAGG_OP_MIN)
TAGG_FOLD_CTOR(min_agg_date, DateTime, TAGG_OP_MIN)
TAGG_FOLD_CTOR(max_agg_u64, uint64_t, TAGG_OP_MAX) TAGG_FOLD_CTOR(max_agg_i64, int64_t, TAGG_OP_MAX) TAGG_FOLD_CTOR(max_agg_f64, double, TAGG_OP_MAX)
TAGG_FOLD_CTOR(max_agg_date, DateTime, TAGG_OP_MAX)
#undef TAGG_FOLD_CTOR

template <bool MULTI>
struct PercentilesAgg {  // percentiles_agg_f64[s] — src/metric/percentile.rs:130-138
    using Fruit = Percentiles;
    Field field;
    mutable uint32_t node = 0;
    void emit(PlanBuilder& pb) const { tagg_node n{}; n.op = TAGG_OP_PERCENTILES; n.kind = TAGG_F64; n.multi = MULTI; n.field_id = field; node = pb.emit(n); }
    Fruit read(const ResultReader& r, uint32_t bucket) const {
        Percentiles p; uint64_t np = 0;
        check(tagg_result_percentiles_len(r.get(), node, bucket, &p.n, &np));
        p.ranks.resize(np);
        std::vector<uint64_t> bits(np);
        check(tagg_result_percentiles_read(r.get(), node, bucket, p.ranks.data(), bits.data(), np));
        p.values.resize(np);
        for (size_t i = 0; i < np; i++) p.values[i] = Kind<double>::from_bits(bits[i]);
        return p;
    }
};
inline PercentilesAgg<false> percentiles_agg_f64(Field f) { return {f}; }
inline PercentilesAgg<true> percentiles_agg_f64s(Field f) { return {f}; }

// (a1, .., an) — src/tuple.rs:5-81 (arity 2..=10)
template <class... A>
struct TupleAgg {
    static_assert(sizeof...(A) >= 2 && sizeof...(A) <= 10, "tuple aggregations have arity 2..=10 (src/tuple.rs:73-81)");
    using Fruit = std::tuple<typename A::Fruit...>;
    std::tuple<A...> members;
    void emit(PlanBuilder& pb) const {
        tagg_node n{}; n.op = TAGG_OP_TUPLE; n.n_children = sizeof...(A); pb.emit(n);
        std::apply([&](auto&... m) { (m.emit(pb), ...); }, members);
    }
    Fruit read(const ResultReader& r, uint32_t bucket) const {
        return std::apply([&](auto&... m) { return Fruit(m.read(r, bucket)...); }, members);
    }
};
template <class... A> TupleAgg<A...> tuple(A... a) { return {std::make_tuple(a...)}; }

// buckets of scope `node` whose parent bucket is `parent`: (key, bucket index)
inline std::vector<std::pair<uint64_t, uint32_t>> children(const ResultReader& r, uint32_t node, uint32_t parent) {
    auto& s = r.scope(node);
    std::vector<std::pair<uint64_t, uint32_t>> out;
    for (size_t i = 0; i < s.keys.size(); i++)
        if (s.parents[i] == parent) out.emplace_back(s.keys[i], (uint32_t)i);
    return out;
}

template <class K, bool MULTI, class Sub>
struct TermsAgg {  // terms_agg_{u64,i64}[s] / filtered_terms_agg_* — src/bucket/terms.rs:185-195,391-401
    using Fruit = Terms<K, typename Sub::Fruit>;
    Field field; Sub sub; std::function<bool(K)> key_filter;
    mutable uint32_t node = 0;
    void emit(PlanBuilder& pb) const {
        tagg_node n{}; n.op = TAGG_OP_TERMS; n.kind = Kind<K>::id; n.multi = MULTI; n.field_id = field; n.n_children = 1; node = pb.emit(n);
        sub.emit(pb);
    }
    Fruit read(const ResultReader& r, uint32_t bucket) const {
        Fruit f;
        for (auto& [bits, child] : children(r, node, bucket)) {
            K key = Kind<K>::from_bits(bits);
            if (key_filter && !key_filter(key)) continue;  // terms.rs:322-330: the filter only decides which buckets exist
            f.res.emplace(key, sub.read(r, child));
        }
        return f;
    }
};
template <class Sub> TermsAgg<uint64_t, false, Sub> terms_agg_u64(Field f, Sub s) { return {f, s, nullptr}; }
template <class Sub> TermsAgg<int64_t, false, Sub> terms_agg_i64(Field f, Sub s) { return {f, s, nullptr}; }
template <class Sub> TermsAgg<uint64_t, true, Sub> terms_agg_u64s(Field f, Sub s) { return {f, s, nullptr}; }
template <class Sub> TermsAgg<int64_t, true, Sub> terms_agg_i64s(Field f, Sub s) { return {f, s, nullptr}; }
template <class Sub, class F> TermsAgg<uint64_t, false, Sub> filtered_terms_agg_u64(Field f, Sub s, F flt) { return {f, s, flt}; }
template <class Sub, class F> TermsAgg<int64_t, false, Sub> filtered_terms_agg_i64(Field f, Sub s, F flt) { return {f, s, flt}; }
template <class Sub, class F> TermsAgg<uint64_t, true, Sub> filtered_terms_agg_u64s(Field f, Sub s, F flt) { return {f, s, flt}; }
template <class Sub, class F> TermsAgg<int64_t, true, Sub> filtered_terms_agg_i64s(Field f, Sub s, F flt) { return {f, s, flt}; }

template <class Sub>
struct HistogramAgg {  // histogram_agg_f64(field, start, interval, sub) — src/bucket/histogram.rs:9-21
    using Fruit = Histogram<typename Sub::Fruit>;
    Field field; double start, interval; Sub sub;
    uint8_t kind = TAGG_F64;  // TAGG_DATE / TAGG_I64: date_histogram_agg below
    mutable uint32_t node = 0;
    void emit(PlanBuilder& pb) const {
        tagg_node n{}; n.op = TAGG_OP_HISTOGRAM; n.kind = kind; n.field_id = field; n.n_children = 1; n.f0 = start; n.f1 = interval; node = pb.emit(n);
        sub.emit(pb);
    }
    Fruit read(const ResultReader& r, uint32_t bucket) const {
        Fruit f; f.start = start; f.interval = interval;
        for (auto& [ord, child] : children(r, node, bucket)) f.bucket_map.emplace(ord, sub.read(r, child));
        return f;
    }
};
template <class Sub> HistogramAgg<Sub> histogram_agg_f64(Field f, double start, double interval, Sub s) { return {f, start, interval, s}; }
// Beyond the reference (its README.md:31-45 TODO list): fixed-width buckets over a date fast field (seconds since the
// epoch), ordinal = floor((t - start) / interval) by the arithmetic of histogram.rs:136-152 on the timestamp
template <class Sub> HistogramAgg<Sub> date_histogram_agg(Field f, uint64_t interval_seconds, Sub s, int64_t start = 0) {
    return {f, (double)start, (double)interval_seconds, s, TAGG_DATE};
}

// cardinality_agg_{u64,i64}[s](field): the EXACT distinct count — the bucket table of terms_agg(field, count_agg()) is the
// distinct set (README.md:36 names an estimate; exact is inside any estimator's tolerance)
template <class K, bool MULTI>
struct CardinalityAgg {
    using Fruit = uint64_t;
    TermsAgg<K, MULTI, CountAgg> inner;
    void emit(PlanBuilder& pb) const { inner.emit(pb); }
    Fruit read(const ResultReader& r, uint32_t bucket) const { return children(r, inner.node, bucket).size(); }
};
inline CardinalityAgg<uint64_t, false> cardinality_agg_u64(Field f) { return {{f, CountAgg(), nullptr}}; }
inline CardinalityAgg<int64_t, false> cardinality_agg_i64(Field f) { return {{f, CountAgg(), nullptr}}; }
inline CardinalityAgg<uint64_t, true> cardinality_agg_u64s(Field f) { return {{f, CountAgg(), nullptr}}; }
inline CardinalityAgg<int64_t, true> cardinality_agg_i64s(Field f) { return {{f, CountAgg(), nullptr}}; }

template <class Sub>
struct FilterAgg {  // filter_agg(&query, sub) — src/filter.rs:8-16; borrows the query like FilterAgg<'q>
    using Fruit = typename Sub::Fruit;
    const Query* query; Sub sub;
    void emit(PlanBuilder& pb) const {
        tagg_node n{}; n.op = TAGG_OP_FILTER; n.n_children = 1; n.aux = (uint32_t)pb.filters.size();
        pb.filters.push_back(query); pb.emit(n);
        sub.emit(pb);
    }
    Fruit read(const ResultReader& r, uint32_t bucket) const { return sub.read(r, bucket); }
};
template <class Sub> FilterAgg<Sub> filter_agg(const Query& q, Sub s) { return {&q, s}; }

// Declarative predicates for post_filter_agg_* (the reference takes a closure, src/post_filter.rs:133-141;
// comparisons lower to an inclusive range on the order-preserving codes)
struct CodeRange { uint64_t lo, hi; };
template <class T> CodeRange gt(T x) { return {Kind<T>::code(x) + 1, ~0ull}; }
template <class T> CodeRange ge(T x) { return {Kind<T>::code(x), ~0ull}; }
template <class T> CodeRange lt(T x) { return {0, Kind<T>::code(x) - 1}; }
template <class T> CodeRange le(T x) { return {0, Kind<T>::code(x)}; }
template <class T> CodeRange eq(T x) { return {Kind<T>::code(x), Kind<T>::code(x)}; }
template <> inline CodeRange gt<double>(double x) { return {f64_to_code(x) + 1, f64_to_code(INFINITY)}; }   // NaN codes lie outside [-inf, +inf]
template <> inline CodeRange ge<double>(double x) { return {f64_to_code(x), f64_to_code(INFINITY)}; }
template <> inline CodeRange lt<double>(double x) { return {f64_to_code(-INFINITY), f64_to_code(x) - 1}; }
template <> inline CodeRange le<double>(double x) { return {f64_to_code(-INFINITY), f64_to_code(x)}; }

template <class T, bool MULTI, class Sub>
struct PostFilterAgg {  // post_filter_agg_{u64,i64,f64}[s] — src/post_filter.rs:303-315
    using Fruit = typename Sub::Fruit;
    Field field; CodeRange range; Sub sub;
    void emit(PlanBuilder& pb) const {
        tagg_node n{}; n.op = TAGG_OP_POST_FILTER; n.kind = Kind<T>::id; n.multi = MULTI; n.field_id = field; n.n_children = 1;
        n.pred = TAGG_PRED_RANGE; n.u0 = range.lo; n.u1 = range.hi; pb.emit(n);
        sub.emit(pb);
    }
    Fruit read(const ResultReader& r, uint32_t bucket) const { return sub.read(r, bucket); }
};
template <class Sub> PostFilterAgg<uint64_t, false, Sub> post_filter_agg_u64(Field f, CodeRange p, Sub s) { return {f, p, s}; }
template <class Sub> PostFilterAgg<int64_t, false, Sub> post_filter_agg_i64(Field f, CodeRange p, Sub s) { return {f, p, s}; }
template <class Sub> PostFilterAgg<double, false, Sub> post_filter_agg_f64(Field f, CodeRange p, Sub s) { return {f, p, s}; }
template <class Sub> PostFilterAgg<uint64_t, true, Sub> post_filter_agg_u64s(Field f, CodeRange p, Sub s) { return {f, p, s}; }
template <class Sub> PostFilterAgg<int64_t, true, Sub> post_filter_agg_i64s(Field f, CodeRange p, Sub s) { return {f, p, s}; }
template <class Sub> PostFilterAgg<double, true, Sub> post_filter_agg_f64s(Field f, CodeRange p, Sub s) { return {f, p, s}; }

// ---- the driver: trait AggSearcher (src/searcher.rs:12-25, 53-101) ---------------------------------------------
enum class Executor { SingleThread, ThreadPool };

class Searcher {
  public:
    Searcher(const Context& ctx, std::vector<const Segment*> segments) : ctx_(ctx), segs_(std::move(segments)) {}
    const std::vector<const Segment*>& segment_readers() const { return segs_; }

    template <class A> typename A::Fruit agg_search(const Query& query, const A& agg) const {
        return agg_search_with_executor(query, agg, Executor::SingleThread);  // searcher.rs:13-17
    }
    template <class A> typename A::Fruit agg_search_with_executor(const Query& query, const A& agg, Executor ex) const {
        PlanBuilder pb;
        agg.emit(pb);  // Agg::prepare -> PreparedAgg
        tagg_plan* plan = nullptr;
        check(tagg_plan_create(ctx_.get(), pb.nodes.data(), (uint32_t)pb.nodes.size(), nullptr, 0, &plan));
        std::unique_ptr<tagg_plan, int (*)(tagg_plan*)> plan_guard(plan, tagg_plan_destroy);
        std::vector<std::vector<tagg_docset>> filters(segs_.size());
        std::vector<tagg_segment_input> inputs(segs_.size());
        for (size_t i = 0; i < segs_.size(); i++) {
            for (auto* fq : pb.filters) filters[i].push_back(fq->docset(*segs_[i], i));
            inputs[i].segment = segs_[i]->get();
            inputs[i].docset = query.docset(*segs_[i], i);
            inputs[i].filters = filters[i].data();
            inputs[i].n_filters = (uint32_t)filters[i].size();
        }
        tagg_result* res = nullptr;
        if (ex == Executor::SingleThread || segs_.empty()) {
            // one harvest threaded through every segment (searcher.rs:66-78)
            check(tagg_execute(plan, inputs.data(), (uint32_t)inputs.size(), &res));
        } else {
            // a fruit per segment, merged in segment order (searcher.rs:79-98)
            for (size_t i = 0; i < inputs.size(); i++) {
                tagg_result* one = nullptr;
                check(tagg_execute(plan, &inputs[i], 1, &one));
                if (!res) res = one;
                else { int rc = tagg_result_merge(res, one); tagg_result_free(one); check(rc); }
            }
        }
        ResultReader reader(res);
        return agg.read(reader, 0);
    }
  private:
    const Context& ctx_;
    std::vector<const Segment*> segs_;
};

}  // namespace tagg

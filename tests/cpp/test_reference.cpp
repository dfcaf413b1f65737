// The reference's unit tests, written against the typed C++ façade (include/tagg.hpp) so they read like the
// originals (file:line cited per test).  Needs a B200; run by tests/test_gpu_cpp_facade.py.
#include <cstdio>
#include <cstdlib>

#include "tagg.hpp"

using namespace tagg;

#define CHECK(cond)                                                                      \
    do {                                                                                 \
        if (!(cond)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); exit(1); } \
    } while (0)

// test_fixtures/src/lib.rs:101-128
struct ProductSchema { Field id = 0, category_id = 1, tag_ids = 2, price = 3, positive_opinion_percent = 4, attr_facets = 5, date_created = 6; };

// test_fixtures/src/lib.rs:30-79 — the five golden documents, one segment, no deletes
static void index_test_products(Segment& seg, const ProductSchema& s) {
    seg.add_column<uint64_t>(s.category_id, {1, 1, 2, 2, 2});
    seg.add_multicolumn<uint64_t>(s.tag_ids, {{111, 112, 211}, {111, 211, 320}, {211, 311}, {320}, {311, 511}});
    seg.add_column<double>(s.price, {9.99, 10.0, 0.5, 50.0, 100.01});
    seg.add_column<uint64_t>(s.positive_opinion_percent, {82, 100, 71, 85, 99});
    // doc 2 has no date: a missing single-valued fast field value reads as 0 (epoch)
    seg.add_column<DateTime>(s.date_created, {{1577836799}, {1577836800}, {0}, {1577833199}, {1577840399}});
}

template <class T> static bool bucket_is(const std::pair<double, const T*>& b, double key, std::optional<T> v) {
    if (b.first != key) return false;
    if (!v) return b.second == nullptr;
    return b.second && *b.second == *v;
}

int main() {
    Context ctx(0);
    ProductSchema schema;
    Segment seg(ctx, 5);
    index_test_products(seg, schema);
    Searcher searcher(ctx, {&seg});
    Segment empty_seg(ctx, 0);
    empty_seg.add_column<uint64_t>(schema.category_id, {});
    Searcher empty(ctx, {&empty_seg});
    AllQuery all;

    for (Executor ex : {Executor::SingleThread, Executor::ThreadPool}) {
        // count.rs:69-80
        CHECK(searcher.agg_search_with_executor(all, count_agg(), ex) == 5);
        // sum.rs:171-192
        CHECK(searcher.agg_search_with_executor(all, sum_agg_u64(schema.positive_opinion_percent), ex) == std::optional<uint64_t>(437));
        CHECK(searcher.agg_search_with_executor(all, sum_agg_f64(schema.price), ex) == std::optional<double>(170.5));
        CHECK(searcher.agg_search_with_executor(all, sum_agg_u64s(schema.tag_ids), ex) == std::optional<uint64_t>(2740));
        // minmax.rs:197-256
        CHECK(searcher.agg_search_with_executor(all, min_agg_u64(schema.positive_opinion_percent), ex) == std::optional<uint64_t>(71));
        CHECK(searcher.agg_search_with_executor(all, min_agg_date(schema.date_created), ex) == std::optional<DateTime>(DateTime{0}));
        CHECK(searcher.agg_search_with_executor(all, min_agg_f64(schema.price), ex) == std::optional<double>(0.5));
        CHECK(searcher.agg_search_with_executor(all, min_agg_u64s(schema.tag_ids), ex) == std::optional<uint64_t>(111));
        CHECK(searcher.agg_search_with_executor(all, max_agg_f64(schema.price), ex) == std::optional<double>(100.01));
        CHECK(searcher.agg_search_with_executor(all, max_agg_u64(schema.positive_opinion_percent), ex) == std::optional<uint64_t>(100));
        CHECK(searcher.agg_search_with_executor(all, max_agg_date(schema.date_created), ex) == std::optional<DateTime>(DateTime{1577840399}));
        CHECK(searcher.agg_search_with_executor(all, max_agg_u64s(schema.tag_ids), ex) == std::optional<uint64_t>(511));
        // tuple.rs:93-110
        auto t = searcher.agg_search_with_executor(all, tuple(count_agg(), min_agg_f64(schema.price), max_agg_f64(schema.price)), ex);
        CHECK(t == std::make_tuple(uint64_t(5), std::optional<double>(0.5), std::optional<double>(100.01)));
        // percentile.rs:190-221
        auto p = searcher.agg_search_with_executor(all, percentiles_agg_f64(schema.price), ex);
        CHECK(p.percentile(0.5) == 10.0 && p.percentile(0.33) == 9.99 && p.percentile(0.7) == 50.0 && p.percentile(0.01) == 0.5 &&
              p.percentile(0.99) == 100.01);
        // terms.rs:473-487
        CHECK(empty.agg_search_with_executor(all, terms_agg_u64(schema.category_id, count_agg()), ex).top_k(10, [](uint64_t b) { return b; }).empty());
        // terms.rs:490-543
        auto cat = searcher.agg_search_with_executor(all, terms_agg_u64(schema.category_id, tuple(count_agg(), min_agg_f64(schema.price))), ex);
        using B = std::tuple<uint64_t, std::optional<double>>;
        CHECK(cat.get(1) && *cat.get(1) == B(2, 9.99));
        CHECK(cat.get(2) && *cat.get(2) == B(3, 0.5));
        auto top = cat.top_k(2, [](const B& b) { return std::get<0>(b); });
        CHECK(top.size() == 2 && top[0].first == 2 && *top[0].second == B(3, 0.5) && top[1].first == 1 && *top[1].second == B(2, 9.99));
        auto top_max_min = cat.top_k(1, [](const B& b) { return *std::get<1>(b); });
        CHECK(top_max_min.size() == 1 && top_max_min[0].first == 1);
        // terms.rs:546-571
        auto even = searcher.agg_search_with_executor(
            all, filtered_terms_agg_u64(schema.category_id, tuple(count_agg(), min_agg_f64(schema.price)), [](uint64_t c) { return c % 2 == 0; }), ex);
        CHECK(even.get(1) == nullptr && even.get(2) && *even.get(2) == B(3, 0.5));
        // histogram.rs:195-223
        auto h = searcher.agg_search_with_executor(all, histogram_agg_f64(schema.price, 0.0, 10.0, count_agg()), ex).buckets();
        CHECK(h.size() == 11);
        const std::optional<uint64_t> none;
        const std::optional<uint64_t> expect0[11] = {2, 1, none, none, none, 1, none, none, none, none, 1};
        for (int i = 0; i < 11; i++) CHECK(bucket_is<uint64_t>(h[i], 10.0 * i, expect0[i]));
        // histogram.rs:226-249
        auto h35 = searcher.agg_search_with_executor(all, histogram_agg_f64(schema.price, 35.0, 10.0, count_agg()), ex).buckets();
        CHECK(h35.size() == 6 && bucket_is<uint64_t>(h35[0], 45.0, 1) && bucket_is<uint64_t>(h35[1], 55.0, none) && bucket_is<uint64_t>(h35[5], 95.0, 1));
        // histogram.rs:252-338
        auto tags = searcher.agg_search_with_executor(
            all, terms_agg_u64s(schema.tag_ids, tuple(count_agg(), histogram_agg_f64(schema.price, 0.0, 10.0, count_agg()))), ex);
        auto top_tags = tags.top_k(3, [](const auto& b) { return std::get<0>(b); });
        CHECK(top_tags.size() == 3 && top_tags[0].first == 211 && std::get<0>(*top_tags[0].second) == 3);
        auto h211 = std::get<1>(*top_tags[0].second).buckets();
        CHECK(h211.size() == 2 && bucket_is<uint64_t>(h211[0], 0.0, 2) && bucket_is<uint64_t>(h211[1], 10.0, 1));
        auto h320 = std::get<1>(*tags.get(320)).buckets();
        CHECK(std::get<0>(*tags.get(320)) == 2 && h320.size() == 5 && bucket_is<uint64_t>(h320[0], 10.0, 1) && bucket_is<uint64_t>(h320[4], 50.0, 1));
        // histogram.rs:341-367
        auto price_query = range_query_f64(schema.price, 10.0, 100.0);
        auto fh = searcher.agg_search_with_executor(all, filter_agg(price_query, histogram_agg_f64(schema.price, 0.0, 10.0, count_agg())), ex).buckets();
        CHECK(fh.size() == 5 && bucket_is<uint64_t>(fh[0], 10.0, 1) && bucket_is<uint64_t>(fh[4], 50.0, 1));
        // filter.rs:137-166
        auto cat1 = term_query_u64(schema.category_id, 1), cat2 = term_query_u64(schema.category_id, 2);
        CHECK(searcher.agg_search_with_executor(all, filter_agg(cat1, count_agg()), ex) == 2);
        CHECK(searcher.agg_search_with_executor(range_query_f64(schema.price, 100.0, 200.0), filter_agg(cat2, count_agg()), ex) == 1);
        // post_filter.rs:330-345
        CHECK(searcher.agg_search_with_executor(all, post_filter_agg_f64(schema.price, gt(5.0), count_agg()), ex) == 4);
    }
    {   // beyond the reference (its TODO list, README.md:31-45): date_histogram and cardinality on the fixture
        auto days = searcher.agg_search(all, date_histogram_agg(schema.date_created, 86400, count_agg()));
        CHECK(days.bucket_map.size() == 3 && days.bucket_map.at(0) == 1 && days.bucket_map.at(18261) == 2 && days.bucket_map.at(18262) == 2);
        auto cats = searcher.agg_search(all, terms_agg_u64(schema.category_id, count_agg()));
        CHECK(searcher.agg_search(all, cardinality_agg_u64(schema.category_id)) == cats.res.size());
        auto tags = searcher.agg_search(all, terms_agg_u64s(schema.tag_ids, count_agg()));
        CHECK(searcher.agg_search(all, cardinality_agg_u64s(schema.tag_ids)) == tags.res.size());
    }
    // FastFieldNotAvailableError (sum.rs:50-55); the reference's histogram panics instead (histogram.rs:81)
    bool threw = false;
    try { searcher.agg_search(all, sum_agg_u64(42)); } catch (const FastFieldNotAvailableError&) { threw = true; }
    CHECK(threw);
    printf("cpp facade: all reference tests passed\n");
    return 0;
}

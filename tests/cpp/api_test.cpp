// C++ mirror of the reference's tests/api_tests.rs against vector-indexer_b200/csrc/api.hpp.
// Built by tests/test_cpp_api.py; "errors" mode runs without a GPU, "full" needs a B200.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../vector-indexer_b200/csrc/api.hpp"

using namespace vector_indexer;

#define EXPECT(c)                                                         \
    do {                                                                  \
        if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

// tests/api_tests.rs:12-25 make_records: value = i*0.01 + j
static std::vector<VectorRecord> make_records(size_t n, size_t dim, uint64_t id0) {
    std::vector<VectorRecord> r;
    for (size_t i = 0; i < n; i++) {
        std::vector<float> v(dim);
        for (size_t j = 0; j < dim; j++) v[j] = (float)i * 0.01f + (float)j;
        r.push_back(VectorRecord{id0 + i, v, std::nullopt});
    }
    return r;
}

template <class F>
static int kind_of(F&& f) {
    try { f(); } catch (const IoError& e) { return e.kind; }
    return 0;
}

int main(int argc, char** argv) {
    std::string mode = argc > 1 ? argv[1] : "errors";
    std::string dir = argc > 2 ? argv[2] : "/tmp/vidx_cpp_api";
    // defaults (tests/api_tests.rs:28-37)
    VectorIndexerConfig cfg(8);
    EXPECT(cfg.default_k == 10 && cfg.default_n_probe == 20 && cfg.max_k == 10000 && cfg.max_n_probe == 10000);
    EXPECT(cfg.index_dir == "index" && cfg.shards_dir == "shards");
    // error cases that need no device (tests/api_tests.rs:252-341)
    EXPECT(kind_of([&] { VectorIndexer(VectorIndexerConfig(8)).build_from_records({}); }) == VIDX_ERR_INVALID_INPUT);
    EXPECT(kind_of([&] {
               auto recs = make_records(4, 8, 0);
               recs[2].values.pop_back();
               VectorIndexer(VectorIndexerConfig(8)).build_from_records(recs);
           }) == VIDX_ERR_INVALID_INPUT);
    EXPECT(kind_of([&] { VectorIndexer::load(VectorIndexerConfig(8).with_index_dir("/nonexistent/x")); }) == VIDX_ERR_NOT_FOUND);
    {
        VectorIndexer v{VectorIndexerConfig(8)};
        auto req = v.search_request(std::vector<float>(7, 0.f));
        EXPECT(req.k == 10 && req.n_probe == 20 && !req.include_vectors);
        EXPECT(kind_of([&] { v.search(req); }) == VIDX_ERR_INVALID_INPUT);  // query dimension mismatch
    }
    if (mode != "full") { printf("OK errors\n"); return 0; }

    // configured dirs, self query returns external_id 42 (tests/api_tests.rs:40-92)
    auto built = VectorIndexer(VectorIndexerConfig(8).with_index_dir(dir + "/index").with_shards_dir(dir + "/shards"))
                     .build_from_records(make_records(300, 8, 42));
    std::vector<float> q(8);
    for (size_t j = 0; j < 8; j++) q[j] = (float)j;
    auto res = built.search(built.search_request(q).with_k(5).with_n_probe(1000).with_include_vectors(true));
    EXPECT(res.size() == 5 && res[0].external_id == 42 && res[0].distance == 0.0f);
    EXPECT(res[0].vector && std::memcmp(res[0].vector->data(), q.data(), 32) == 0);
    for (size_t t = 1; t < res.size(); t++) EXPECT(res[t].distance >= res[t - 1].distance);
    // k clamped to max_k, n_probe clamped (tests/api_tests.rs:95-197)
    EXPECT(built.search(built.search_request(q).with_k(100000).with_n_probe(100000)).size() == 300);
    EXPECT(kind_of([&] { built.search(built.search_request(q).with_k(0)); }) == VIDX_ERR_INVALID_INPUT);
    // reload from the files just written
    auto loaded = VectorIndexer::load(VectorIndexerConfig(8).with_index_dir(dir + "/index").with_shards_dir(dir + "/shards"));
    auto res2 = loaded.search(loaded.search_request(q).with_k(5).with_n_probe(1000));
    EXPECT(res2.size() == 5 && res2[0].external_id == 42);
    for (size_t t = 0; t < 5; t++) EXPECT(res2[t].external_id == res[t].external_id && res2[t].distance == res[t].distance);
    printf("OK full\n");
    return 0;
}

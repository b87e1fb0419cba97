// api.hpp -- header-only C++17 mirror of the reference's public API (src/api.rs) over the
// C ABI of include/vidx_b200.h.  Same names, argument meaning and error behaviour:
// errors are std::system_error-like exceptions carrying the io::ErrorKind-equivalent code.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/vidx_b200.h"

namespace vector_indexer {

// std::io::Error
struct IoError : std::runtime_error {
    int kind;  // VIDX_ERR_* (= io::ErrorKind)
    IoError(int k, const std::string& m) : std::runtime_error(m), kind(k) {}
};

// src/api.rs:9-54
struct VectorIndexerConfig {
    uint32_t dimension;
    std::string index_dir = "index";
    std::string shards_dir = "shards";
    size_t default_k = 10, default_n_probe = 20, max_k = 10000, max_n_probe = 10000;
    int device = 0;  // (addition) CUDA ordinal
    explicit VectorIndexerConfig(uint32_t dim) : dimension(dim) {}
    VectorIndexerConfig with_index_dir(std::string d) && { index_dir = std::move(d); return std::move(*this); }
    VectorIndexerConfig with_shards_dir(std::string d) && { shards_dir = std::move(d); return std::move(*this); }
};

// src/api.rs:57-62
struct VectorRecord {
    uint64_t external_id;
    std::vector<float> values;
    std::optional<uint64_t> timestamp;  // nullopt => now
};

// src/api.rs:64-87
struct SearchRequest {
    std::vector<float> query;
    bool include_vectors = false;
    size_t k = 10, n_probe = 20;
    SearchRequest with_k(size_t v) && { k = v; return std::move(*this); }
    SearchRequest with_n_probe(size_t v) && { n_probe = v; return std::move(*this); }
    SearchRequest with_include_vectors(bool v) && { include_vectors = v; return std::move(*this); }
};

// src/api.rs:89-94
struct SearchResult {
    uint64_t external_id;
    float distance;  // squared L2
    std::optional<std::vector<float>> vector;
};

// src/api.rs:96-237
class VectorIndexer {
  public:
    explicit VectorIndexer(VectorIndexerConfig cfg) : cfg_(std::move(cfg)) {
        check(vidx_create(cfg_.dimension, cfg_.device, &h_));
        check(vidx_set_limits(h_, cfg_.default_k, cfg_.default_n_probe, cfg_.max_k, cfg_.max_n_probe));
    }
    ~VectorIndexer() { if (h_) vidx_free(h_); }
    VectorIndexer(VectorIndexer&& o) noexcept : cfg_(std::move(o.cfg_)), h_(o.h_) { o.h_ = nullptr; }
    VectorIndexer(const VectorIndexer&) = delete;
    VectorIndexer& operator=(const VectorIndexer&) = delete;

    // VectorIndexer::load (src/api.rs:109-112)
    static VectorIndexer load(VectorIndexerConfig cfg) {
        VectorIndexer v(std::move(cfg));
        check(vidx_load(v.h_, v.cfg_.index_dir.c_str(), v.cfg_.shards_dir.c_str()));
        return v;
    }

    // build_from_records (src/api.rs:115-146): validates, trains with seed 42, persists
    VectorIndexer build_from_records(std::vector<VectorRecord> records) && {
        if (records.empty()) throw IoError(VIDX_ERR_INVALID_INPUT, "no vectors provided");
        const size_t dim = cfg_.dimension;
        std::vector<float> vals;
        std::vector<uint64_t> ids, ts;
        vals.reserve(records.size() * dim);
        for (size_t i = 0; i < records.size(); i++) {
            const auto& r = records[i];
            if (r.values.size() != dim)
                throw IoError(VIDX_ERR_INVALID_INPUT, "vector dimension mismatch at index " + std::to_string(i) + ": expected " +
                                                           std::to_string(dim) + ", got " + std::to_string(r.values.size()));
            vals.insert(vals.end(), r.values.begin(), r.values.end());
            ids.push_back(r.external_id);
            ts.push_back(r.timestamp.value_or(0));
        }
        check(vidx_build(h_, vals.data(), ids.data(), ts.data(), records.size(), 42, 0, 0));
        check(vidx_save(h_, cfg_.index_dir.c_str(), cfg_.shards_dir.c_str()));
        return std::move(*this);
    }

    // search (src/api.rs:188-222)
    std::vector<SearchResult> search(const SearchRequest& req) const {
        if (req.query.size() != cfg_.dimension)
            throw IoError(VIDX_ERR_INVALID_INPUT, "query dimension mismatch: expected " + std::to_string(cfg_.dimension) +
                                                       ", got " + std::to_string(req.query.size()));
        const size_t k = req.k < cfg_.max_k ? req.k : cfg_.max_k;
        if (k == 0 || req.n_probe == 0) throw IoError(VIDX_ERR_INVALID_INPUT, "k and n_probe must be greater than 0");
        std::vector<float> d(k), v;
        std::vector<int64_t> ids(k);
        if (req.include_vectors) {
            v.resize(k * cfg_.dimension);
            check(vidx_search_with_vectors(h_, req.query.data(), 1, k, req.n_probe, d.data(), ids.data(), v.data()));
        } else {
            check(vidx_search(h_, req.query.data(), 1, k, req.n_probe, d.data(), ids.data()));
        }
        std::vector<SearchResult> out;
        for (size_t t = 0; t < k && ids[t] >= 0; t++) {
            SearchResult r{(uint64_t)ids[t], d[t], std::nullopt};
            if (req.include_vectors) r.vector = std::vector<float>(v.begin() + t * cfg_.dimension, v.begin() + (t + 1) * cfg_.dimension);
            out.push_back(std::move(r));
        }
        return out;
    }

    // search_request (src/api.rs:225-232)
    SearchRequest search_request(std::vector<float> query) const {
        SearchRequest r;
        r.query = std::move(query);
        r.k = cfg_.default_k;
        r.n_probe = cfg_.default_n_probe;
        return r;
    }
    const VectorIndexerConfig& config() const { return cfg_; }
    vidx_index* handle() const { return h_; }

  private:
    static void check(int rc) {
        if (rc != VIDX_OK) throw IoError(rc, vidx_last_error());
    }
    VectorIndexerConfig cfg_;
    vidx_index* h_ = nullptr;
};

}  // namespace vector_indexer

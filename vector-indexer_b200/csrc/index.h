// index.h -- host-side state of the device-resident IVF index.
#pragma once
#include <string>
#include <vector>

#include "../../include/vidx_b200.h"
#include "common.cuh"
#include "scan_tc.h"

namespace vidx {

extern thread_local std::string t_last_error;

struct Index {
    // configuration (VectorIndexerConfig, src/api.rs:9-54)
    uint32_t dim = 0;
    int device = 0;
    uint64_t default_k = 10, default_n_probe = 20, max_k = 10000, max_n_probe = 10000;
    uint64_t seed = 42;

    // training result (unfiltered numbering, src/ivf_index.rs:70-109)
    bool trained = false, built = false;
    uint64_t k_trained = 0, num_shards = 0, train_iters = 0;
    std::vector<float> train_centroids;   // k_trained x dim
    std::vector<uint32_t> train_labels;   // per vector
    std::vector<uint32_t> super_labels;   // per trained centroid -> shard

    // lists after the empty-list filter (src/ivf_index.rs:122-164)
    uint64_t ntotal = 0, nlist = 0;
    std::vector<float> centroids;         // nlist x dim
    std::vector<uint32_t> c2shard;        // centroids_to_shard
    std::vector<uint32_t> list_len;
    std::vector<uint32_t> old_to_new;     // trained id -> list id (kNoRow if dropped)
    std::vector<uint64_t> list_goff;      // first group of each list
    std::vector<uint32_t> list_seg_off_all;  // CSR: segments of each list
    std::vector<SegDesc> segs;
    std::vector<uint32_t> row_src;        // row -> internal id (build order), kNoRow = padding
    std::vector<uint64_t> ext_ids, timestamps;  // per internal id
    std::vector<std::string> load_warnings;  // shards skipped by vidx_load (ivf_index.rs:254 drops failed shard reads)
    std::vector<uint64_t> internal_ids;   // loaded indexes only: build position -> VectorMeta.id (shards.rs:45-51)

    // partition (multi-GPU): lists owned by this rank keep their segment range
    int part_rank = 0, part_world = 1;
    int part_mode = 0;                    // 0 auto, 1 shards, 2 segment ranges of every list
    bool part_by_ranges = false;          // what apply_partition chose
    uint64_t owned_vectors = 0;
    std::vector<uint2> list_seg_part;     // per list (first segment, end segment); empty if not owned
    std::vector<uint64_t> seg_prefix;     // prefix sums of per-list segment counts, largest first
    std::vector<uint64_t> tile_prefix;    // prefix sums of per-list 128-vector tile counts (owned part), largest first

    // device store
    cudaStream_t stream = nullptr;
    cudaEvent_t events[10] = {};
    uint32_t ncgroups = 0;
    DevBuf d_vecs, d_cents, d_row_ext, d_segs, d_list_seg, d_list_g0, d_list_ng, d_list_len, d_vnorm, d_vecs16;
    float vn_max = 0.0f;   // max |v|^2 over the stored rows (bound for the tensor-core filter)
    float vmax = 0.0f;     // max |component| over the stored rows
    int tc_sv = 0, tc_g = 0;  // fp16 shadow store: vectors scaled by 2^sv, norm terms by 2^(2sv-g)
    bool tc_ok = false;    // the shadow store exists (finite data of sane magnitude)
    // The centroid table as a one-list index of its own: coarse quantization = the same tensor-core filter + exact
    // re-check, top-n_probe by (distance, list id).
    struct CoarseTable {
        DevBuf vecs16, vnorm, list_g0, list_ng, list_len, list_seg;
        float vn_max = 0.0f, vmax = 0.0f;
        int sv = 0, g = 0;
        bool ok = false;
    } ctab;
    int coarse_mode = 0;   // 0 = tensor-core filter when n_probe <= 32 and nlist is large enough, 1 = exact kernels only, 2 = filter whenever possible
    void coarse_tc(const float4* xq4, uint32_t nqb, uint32_t np, uint32_t* d_probes, float* d_probe_dist, cudaStream_t st);
    int scan_mode = 0;     // 0 = tensor-core filter when the shape allows (bounds pass first when a query visits few tiles), 1 = exact kernels
                           // only, 2 = filter with seeding pass only, 3 = filter with bounds pass whenever its minima fit
    DevBuf io_xq, io_D, io_I, io_rows, io_V;
    struct Workspace;
    Workspace* ws = nullptr;

    // measurement
    bool profiling = false;
    double st_ms[6] = {};
    vidx_search_stats stats{};

    int dq() const { return (int)((dim + 3) / 4); }
    void ensure_device();
    void delete_workspace();
    void train_on_device(const float* d_data, uint64_t n, uint64_t seed, uint64_t nlist_override, uint64_t iters_override,
                         DevBuf& d_labels);
    void build_lists(const float* d_data, uint64_t n, const uint32_t* labels, const uint64_t* ext, const uint64_t* ts,
                     const float* cents_all, uint64_t k, const uint32_t* shard_of_centroid, bool keep_empty = false);
    std::vector<int32_t> shard_owners(int world) const;
    void apply_partition();
    void search_device(const float* d_xq, uint64_t nq, uint64_t k, uint64_t nprobe, float* d_D, int64_t* d_I, uint32_t* d_rows,
                       cudaStream_t st, uint32_t* d_probe_out, float* d_probe_dist_out);
    ~Index();
};


std::vector<int32_t> partition_shards_public(const std::vector<uint64_t>& load, int world);

// What vidx_load reads from index.bin + the shard files.
struct LoadedIndex {
    uint32_t dim = 0;
    uint64_t nlist = 0, num_shards = 0;
    std::vector<float> centroids;
    std::vector<uint32_t> c2shard;
    std::vector<std::vector<float>> list_vectors;   // per list, len x dim
    std::vector<std::vector<uint64_t>> list_meta;   // per list, (id, external_id, timestamp) per vector
    std::vector<std::string> skipped_shards;
};
void save_index(const Index& ix, const std::vector<float>& host_vectors, const std::string& index_dir,
                const std::string& shards_dir);
void load_index_files(const std::string& index_dir, const std::string& shards_dir, uint32_t expect_dim, LoadedIndex& out);
// vector files (src/utils.rs:34-107): concatenated bincode batches of (id, values, metadata)
void read_vector_file(const std::string& path, std::vector<uint64_t>& ids, std::vector<uint64_t>& lens, std::vector<float>& values,
                      std::vector<uint64_t>& meta);
void write_vector_file(const std::string& path, const float* data, const uint64_t* ids, const uint64_t* meta, uint64_t n, uint64_t dim,
                       uint64_t batch);

}  // namespace vidx

// index.h -- host-side state of the device-resident IVF index.
#pragma once
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vidx_b200.h"
#include "common.cuh"
#include "scan_tc.h"

namespace vidx {

extern thread_local std::string t_last_error;

struct SearchCtx;  // per in-flight search: stream, events, workspace (index.cu)
struct Comm;       // NCCL communicator of a multi-GPU index (comm.cu)

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

    // lists after the empty-list filter (src/ivf_index.rs:122-164).  Everything in this block describes the WHOLE
    // index and is identical on every rank of a partitioned index ("global" groups / segments / rows).
    uint64_t ntotal = 0, nlist = 0;
    std::vector<float> centroids;         // nlist x dim
    std::vector<uint32_t> c2shard;        // centroids_to_shard
    std::vector<uint32_t> list_len;
    std::vector<uint32_t> old_to_new;     // trained id -> list id (kNoRow if dropped)
    std::vector<uint64_t> list_goff;      // first global group of each list
    std::vector<uint32_t> list_seg_off_all;  // CSR: global segments of each list
    std::vector<SegDesc> segs;            // global segment table (g0 = global group)

    // What sits in HBM.  A handle whose partition was set BEFORE build / load keeps only the part its rank owns
    // (resident_partial): rows are then numbered locally, list by list over the resident segment ranges.
    bool part_pending = false;            // vidx_set_partition came before build / load
    bool resident_partial = false;
    std::vector<uint2> res_seg;           // per list: resident range of global segment ids
    std::vector<uint32_t> res_g0;         // per list: first LOCAL group of the resident range
    uint64_t res_groups = 0;              // groups in HBM
    uint64_t resident_vectors = 0;
    std::vector<uint32_t> row_src;        // local row -> build position, kNoRow = padding
    std::vector<uint64_t> ext_ids, timestamps;  // per build position
    std::vector<std::string> load_warnings;  // shards skipped by vidx_load (ivf_index.rs:254 drops failed shard reads)
    std::vector<uint64_t> internal_ids;   // loaded indexes only: build position -> VectorMeta.id (shards.rs:45-51)

    // partition (multi-GPU): the part of the index this rank scans
    int part_rank = 0, part_world = 1;
    int part_mode = 0;                    // 0 auto, 1 shards, 2 segment ranges of every list
    bool part_by_ranges = false;          // what plan_partition chose
    uint64_t owned_vectors = 0;
    std::vector<uint2> list_seg_part;     // per list (first segment, end segment), global ids; empty if not owned
    std::vector<uint64_t> seg_prefix;     // prefix sums of per-list segment counts, largest first
    std::vector<uint64_t> tile_prefix;    // prefix sums of per-list 128-vector tile counts (owned part), largest first
    double mean_probe_tiles = 0.0;        // owned tiles of the list a random vector belongs to (what a probe is expected to cost)

    // device store
    cudaStream_t stream = nullptr;        // build / load / save stream
    uint32_t ncgroups = 0;
    DevBuf d_vecs, d_cents, d_row_ext, d_segs, d_list_seg, d_list_g0, d_list_ng, d_list_len, d_vnorm, d_vecs16, d_list_rowdelta,
        d_shadow_perm, d_vnorm32, d_gmin;  // shadow row -> store row (norm-sorted segments), fp32 norm terms, their group minima
    float vn_max = 0.0f;   // max |v|^2 over the stored rows (bound for the tensor-core filter)
    float vmax = 0.0f;     // max |component| over the stored rows
    int tc_sv = 0, tc_g = 0;  // fp16 shadow store: vectors scaled by 2^sv, norm terms by 2^(2sv-g)
    bool tc_ok = false;    // the shadow store exists (finite data of sane magnitude)
    CUtensorMap shadow_tmap{};  // the shadow store as a tiled-TMA tensor (CTA-pair scan kernel)
    // The centroid table as a one-list index of its own: coarse quantization = the same tensor-core filter + exact
    // re-check, top-n_probe by (distance, list id).
    struct CoarseTable {
        DevBuf vecs16, vnorm, list_g0, list_ng, list_len, list_seg;
        float vn_max = 0.0f, vmax = 0.0f;
        int sv = 0, g = 0;
        bool ok = false;
    } ctab;
    int coarse_mode = 0;   // 0 = tensor-core filter when n_probe <= 32 and nlist is large enough, 1 = exact kernels only, 2 = filter whenever possible
    int scan_mode = 0;     // 0 = tensor-core filter when the shape allows (bounds pass first when a query visits few tiles), 1 = exact kernels
                           // only, 2 = filter with seeding pass only, 3 = filter with bounds pass whenever its minima fit

    // searches in flight: each takes a context (stream-ordered workspace) from the pool, so concurrent calls on one
    // handle never share scratch memory (tests/ivf_index_tests.rs:768-807 searches from several threads)
    std::mutex pool_mu;
    std::vector<SearchCtx*> pool;
    SearchCtx* acquire_ctx();
    void release_ctx(SearchCtx* c, cudaStream_t last_stream);

    Comm* comm = nullptr;  // vidx_comm_init
    uint64_t epoch = 0;    // bumped by every call that takes the handle exclusively (build, load, set_*): cached search graphs of an
                           // older epoch are never replayed

    // measurement
    bool profiling = false;
    vidx_search_stats stats{};

    int dq() const { return (int)((dim + 3) / 4); }
    void ensure_device();
    void delete_workspace();
    void train_on_device(const float* d_data, uint64_t n, uint64_t seed, uint64_t nlist_override, uint64_t iters_override,
                         DevBuf& d_labels);
    void build_lists(const float* d_data, uint64_t n, const uint32_t* labels, const uint64_t* ext, const uint64_t* ts,
                     const float* cents_all, uint64_t k, const uint32_t* shard_of_centroid);
    // the pieces of build_lists, shared with vidx_load
    void layout_lists();                  // list_len -> list_goff, segs, list_seg_off_all
    void plan_partition();                // part_rank / part_world / part_mode -> list_seg_part, part_by_ranges (host only)
    void plan_residency();                // list_seg_part -> res_seg, res_g0, res_groups
    uint32_t local_row(uint64_t list, uint32_t j) const;  // j-th vector of a list -> local row, kNoRow if not resident
    uint32_t local_group_of_seg(uint64_t list, uint32_t seg) const;
    bool list_fully_resident(uint64_t list) const;
    void finish_store(const float* d_data);  // row_src (local) + d_data -> device store, shadow store, partition tables
    void upload_partition();
    std::vector<int32_t> shard_owners(int world) const;
    void apply_partition();
    uint64_t resident_bytes() const;
    void search_device(SearchCtx& c, const float* d_xq, uint64_t nq, uint64_t k, uint64_t nprobe, float* d_D, int64_t* d_I,
                       uint32_t* d_rows, cudaStream_t st, uint32_t* d_probe_out, float* d_probe_dist_out,
                       const uint32_t* d_probes_in = nullptr, unsigned long long* d_keys_out = nullptr, Comm* bounds_comm = nullptr);
    void coarse_tc(SearchCtx& c, const float4* xq4, uint32_t nqb, uint32_t np, uint32_t* d_probes, float* d_probe_dist,
                   cudaStream_t st);
    ~Index();
};

std::vector<int32_t> partition_shards_public(const std::vector<uint64_t>& load, int world);
bool choose_partition_split(const uint32_t* list_len, const uint32_t* c2shard, uint64_t nlist, uint64_t num_shards, int world,
                            int mode, std::vector<int32_t>& owner);

// What vidx_load reads: index.bin and the header + centroid index of every shard file first (list sizes and block
// positions: the partition is decided on these), then only the vector ranges that become resident on this rank.
struct LoadedMeta {
    uint32_t dim = 0;
    uint64_t nlist = 0, num_shards = 0;
    std::vector<float> centroids;
    std::vector<uint32_t> c2shard;
    std::vector<uint32_t> list_len;        // 0 for lists of skipped shards
    std::vector<uint32_t> list_file_shard; // shard file that holds the list's block
    std::vector<uint64_t> list_block_off;  // absolute offset of the block in that file
    std::vector<std::string> skipped_shards;
};
void save_index(const Index& ix, const std::vector<float>& host_vectors, const std::string& index_dir,
                const std::string& shards_dir);
void load_index_meta(const std::string& index_dir, const std::string& shards_dir, LoadedMeta& out);
// index.bin alone (the bincode / ndarray-serde framing of IvfIndex, ivf_index.rs:36-41, :274-316)
void write_index_bin(const std::string& index_dir, const float* centroids, const uint32_t* c2shard, uint64_t nlist, uint32_t D);
void read_index_bin(const std::string& index_dir, LoadedMeta& out);  // fills dim, nlist, num_shards, centroids, c2shard
// Vectors [v0[l], v1[l]) of every list, appended list by list: data (n x dim) and meta (id, external_id, timestamp per vector).
void load_list_ranges(const std::string& shards_dir, const LoadedMeta& m, const std::vector<uint32_t>& v0,
                      const std::vector<uint32_t>& v1, std::vector<float>& data, std::vector<uint64_t>& meta);
// vector files (src/utils.rs:34-107): concatenated bincode batches of (id, values, metadata)
void read_vector_file(const std::string& path, std::vector<uint64_t>& ids, std::vector<uint64_t>& lens, std::vector<float>& values,
                      std::vector<uint64_t>& meta);
void write_vector_file(const std::string& path, const float* data, const uint64_t* ids, const uint64_t* meta, uint64_t n, uint64_t dim,
                       uint64_t batch);

// ---- NCCL (comm.cu): loaded with dlopen when a communicator is first asked for ---------------------------------------
void comm_unique_id(uint8_t out[128]);
Comm* comm_create(int device, int rank, int world, const uint8_t id[128]);
void comm_destroy(Comm* c);
int comm_rank(const Comm* c);
int comm_world(const Comm* c);
void comm_all_gather(Comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t st);
const char* comm_version(const Comm* c);

}  // namespace vidx

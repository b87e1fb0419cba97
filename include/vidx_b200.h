/* vidx_b200.h -- C ABI of libvidx_b200.so: the B200 (sm_100a) implementation of the
 * IVF search + k-means hot path of the Rust crate `vector_indexer`
 * (NirajNair/vector-indexer).  Plain pointers and sizes only; no torch / C++ types.
 *
 * Every entry point names the reference interface it replaces (file:line relative
 * to the reference tree).  Return value: 0 on success, otherwise one of the
 * VIDX_ERR_* codes, which map 1:1 onto the std::io::ErrorKind values the reference
 * returns for the same condition.  vidx_last_error() gives the message.
 *
 * Threading: a handle may be searched from several threads at once -- every search
 * takes its own stream-ordered context (scratch buffers + stream) from a pool in the
 * handle, as the reference shares one index between OS threads
 * (tests/ivf_index_tests.rs:768-807).  vidx_search_device calls on different streams
 * are ordered against each other only through those contexts: results are ready when
 * the stream given to the call reaches that point.  build / load / set_* are exclusive
 * (they wait for running searches and block new ones).
 * Repeated searches: the second call with the same arguments (pointers, sizes, k, n_probe) on
 * a context is captured into a CUDA graph and later ones replay it with one launch
 * (DESIGN.md 4.4); the buffers may hold new contents, results are those of a plain call.
 * A caller stream that is itself being captured is left alone.  VIDX_GRAPH=0 disables it.
 * vidx_search_multi replays too when the index is replicated (partition world 1).
 * There is NO CPU fallback: every compute entry point fails with VIDX_ERR_CUDA when
 * no sm_100 device is usable.
 */
#ifndef VIDX_B200_H
#define VIDX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIDX_OK 0
#define VIDX_ERR_INVALID_INPUT 1 /* io::ErrorKind::InvalidInput */
#define VIDX_ERR_NOT_FOUND 2     /* io::ErrorKind::NotFound     */
#define VIDX_ERR_INVALID_DATA 3  /* io::ErrorKind::InvalidData  */
#define VIDX_ERR_OTHER 4         /* io::ErrorKind::Other        */
#define VIDX_ERR_CUDA 5          /* device / driver failure (no reference counterpart) */
#define VIDX_ERR_UNSUPPORTED 6   /* valid in the reference, outside this build's limits */

typedef struct vidx_index vidx_index;

/* ---- lifecycle -------------------------------------------------------------------- */

/* VectorIndexer::new(VectorIndexerConfig::new(dimension))          src/api.rs:33-43, :103-106
 * `device` is the CUDA ordinal this handle lives on. */
int vidx_create(uint32_t dimension, int device, vidx_index** out);
void vidx_free(vidx_index* idx);
/* Message of the last failing call on this thread (never NULL). */
const char* vidx_last_error(void);

/* VectorIndexerConfig defaults / clamps                             src/api.rs:38-41, :189-190 */
int vidx_set_limits(vidx_index* idx, uint64_t default_k, uint64_t default_n_probe, uint64_t max_k,
                    uint64_t max_n_probe);

/* ---- build = train + add ---------------------------------------------------------- */

/* VectorIndexer::build_from_records -> IvfIndex::fit_with_paths    src/api.rs:115-146, src/ivf_index.rs:58-177
 * data: n x dimension row-major fp32 (VectorStore::get_vectors, src/vector_store.rs:48-58).
 * ext_ids: n external ids, NULL => row index (bindings/python/src/lib.rs:256-262).
 * timestamps: n values, NULL or 0 => now (src/vector_store.rs:36-40).
 * seed: 42 in the reference (src/api.rs:143).  nlist / max_iters: 0 => the reference
 * heuristics (src/utils.rs:9-26); non-zero overrides them (BASELINE configs fix nlist).
 * n == 0 => VIDX_ERR_INVALID_INPUT ("no vectors provided", src/api.rs:116-118). */
int vidx_build(vidx_index* idx, const float* data, const uint64_t* ext_ids, const uint64_t* timestamps, uint64_t n,
               uint64_t seed, uint64_t nlist, uint64_t max_iters);

/* The same with `d_data` already in DEVICE memory of the handle's GPU (n x dimension row-major fp32): no host copy of the
 * data set is needed, which is what a 100M-vector build on a box with less host RAM than HBM requires.  ext_ids / timestamps
 * stay host arrays (or NULL). */
int vidx_build_device(vidx_index* idx, const float* d_data, const uint64_t* ext_ids, const uint64_t* timestamps, uint64_t n,
                      uint64_t seed, uint64_t nlist, uint64_t max_iters);

/* VectorIndexer::build_from_vector_file (src/api.rs:149-186): the file is a concatenation of bincode-2
 * `standard()` batches of (id u64, values Vec<f32>, metadata u64) (src/utils.rs:34-107); decoding stops
 * silently at the first batch that does not decode, as read_vectors_from_file does.  An empty file or a
 * record whose length differs from the index dimension is VIDX_ERR_INVALID_INPUT with the reference's
 * messages; file id = external id, metadata = timestamp (0 -> now, vector_store.rs:36-40).
 * vidx_vector_file_read: data == NULL only counts (n_out); otherwise fills data[n][dim], ids[n], meta[n]
 * (cap = rows the buffers hold).  vidx_vector_file_write writes the same framing (batches of `batch`
 * records, 0 = 1000; ids == NULL: 0..n-1, meta == NULL: 0). */
int vidx_build_from_vector_file(vidx_index* idx, const char* vector_file, uint64_t seed, uint64_t nlist, uint64_t max_iters);
int vidx_vector_file_read(const char* vector_file, uint64_t dim, uint64_t cap, float* data, uint64_t* ids, uint64_t* meta,
                          uint64_t* n_out);
int vidx_vector_file_write(const char* vector_file, const float* data, const uint64_t* ids, const uint64_t* meta, uint64_t n,
                           uint64_t dim, uint64_t batch);

/* The two halves of fit_with_paths, exposed separately (north_star "build/train, add"):
 * train  = mini-batch k-means + super-centroid shard assignment    src/ivf_index.rs:59-77, :103-109
 * add    = list assignment by the trained centroids, list build,
 *          empty-list filter + renumbering                          src/ivf_index.rs:79-164
 * vidx_train keeps the final labels of the training vectors; vidx_add with the same
 * `data` pointer contents reproduces vidx_build exactly. */
int vidx_train(vidx_index* idx, const float* data, uint64_t n, uint64_t seed, uint64_t nlist, uint64_t max_iters);
int vidx_add(vidx_index* idx, const float* data, const uint64_t* ext_ids, const uint64_t* timestamps, uint64_t n);

/* Build the device-resident lists from caller-supplied centroids and labels (no
 * training); the list-build half of src/ivf_index.rs:79-164 on its own. */
int vidx_build_from_labels(vidx_index* idx, const float* data, const uint64_t* ext_ids, uint64_t n,
                           const float* centroids, uint64_t k, const uint64_t* labels, const uint64_t* centroid_shard);

/* ---- search ----------------------------------------------------------------------- */

/* PyVectorIndex.search_blocking(xq, k, n_probe) -> (D, I)           bindings/python/src/lib.rs:123-203
 *   = for each row: VectorIndexer::search -> IvfIndex::search_with_paths
 *                                                                   src/api.rs:188-222, src/ivf_index.rs:190-267
 * xq: nq x dimension fp32 HOST memory.  D: nq x k squared-L2 distances, ascending,
 * padded with +inf.  I: nq x k external ids as i64, padded with -1.
 * k == 0 or n_probe == 0 => VIDX_ERR_INVALID_INPUT (src/ivf_index.rs:197-202);
 * k / n_probe above max_k / max_n_probe are clamped (src/api.rs:189-190); the output
 * stride stays the caller's k. */
int vidx_search(vidx_index* idx, const float* xq, uint64_t nq, uint64_t k, uint64_t n_probe, float* D, int64_t* I);

/* Same with DEVICE pointers (queries already in HBM, results left in HBM) on the given
 * CUDA stream (cudaStream_t cast to void*, NULL = the handle's own stream).  Returns
 * after enqueueing; results are ready when the stream reaches this point. */
int vidx_search_device(vidx_index* idx, const float* d_xq, uint64_t nq, uint64_t k, uint64_t n_probe, float* d_D,
                       int64_t* d_I, void* stream);

/* SearchRequest::with_include_vectors(true)                          src/api.rs:83-86, :213-217
 * As vidx_search, additionally V: nq x k x dimension fp32 payloads (rows of missing
 * results are zero). */
int vidx_search_with_vectors(vidx_index* idx, const float* xq, uint64_t nq, uint64_t k, uint64_t n_probe, float* D,
                             int64_t* I, float* V);

/* Coarse quantization only: the n_probe nearest lists per query, in the order the
 * reference's stable sort yields (src/ivf_index.rs:205-220).  lists: nq x n_probe
 * (padded with UINT32_MAX when n_probe > nlist), dists likewise (+inf). */
int vidx_coarse_probes(vidx_index* idx, const float* xq, uint64_t nq, uint64_t n_probe, uint32_t* lists,
                       float* dists);

/* ---- introspection (what IvfIndex / Shard hold after fit)   src/ivf_index.rs:36-41, :174-176 */
uint32_t vidx_dimension(const vidx_index* idx);
uint64_t vidx_ntotal(const vidx_index* idx);
uint64_t vidx_nlist(const vidx_index* idx);      /* non-empty lists (src/ivf_index.rs:122-146) */
uint64_t vidx_num_shards(const vidx_index* idx); /* ceil(sqrt(k)) (src/ivf_index.rs:104) */
uint64_t vidx_k_trained(const vidx_index* idx);  /* k before the empty-list filter */
int vidx_get_centroids(const vidx_index* idx, float* out /* nlist x dim */);
int vidx_get_centroids_to_shard(const vidx_index* idx, uint64_t* out /* nlist */);
int vidx_get_list_sizes(const vidx_index* idx, uint64_t* out /* nlist */);
/* internal ids (= build row index, src/vector_store.rs:33) of one list, ascending */
int vidx_get_list_members(const vidx_index* idx, uint64_t list, uint64_t* out);
int vidx_get_train_labels(const vidx_index* idx, uint64_t* out /* n, unfiltered numbering */);
int vidx_get_train_centroids(const vidx_index* idx, float* out /* k_trained x dim */);
/* What this handle keeps in HBM: vectors, and bytes of the vector store (fp32 rows + fp16 shadow + norm terms + ids).
 * On a rank of a partitioned index (vidx_set_partition before build / load) both are ~1/world of the whole. */
uint64_t vidx_resident_vectors(const vidx_index* idx);
uint64_t vidx_resident_bytes(const vidx_index* idx);

/* ---- k-means (the functions src/kmeans.rs exports) -------------------------------- */

/* run_kmeans_mini_batch(data, k, max_iters, early_stop_threshold, seed)   src/kmeans.rs:64-150
 * tol < 0 => None => 1e-4.  n == 0 => VIDX_ERR_INVALID_INPUT (src/kmeans.rs:72-77).
 * out_labels are usize in the reference => u64 here.  iters_run may be NULL. */
int vidx_kmeans_mini_batch(int device, const float* data, uint64_t n, uint64_t dim, uint64_t k, uint64_t max_iters,
                           float tol, uint64_t seed, float* out_centroids, uint64_t* out_labels, uint64_t* iters_run);
/* run_kmeans_parallel (full-batch Lloyd)                                   src/kmeans.rs:15-60 */
int vidx_kmeans_parallel(int device, const float* data, uint64_t n, uint64_t dim, uint64_t k, uint64_t max_iters,
                         float tol, uint64_t seed, float* out_centroids, uint64_t* out_labels, uint64_t* iters_run);
/* assign_points_simd_parallel: brute force for k <= 100, hierarchical above
 *                                                                          src/kmeans.rs:445-581 */
int vidx_assign_points(int device, const float* data, uint64_t n, uint64_t dim, const float* centroids, uint64_t k,
                       uint64_t seed, uint64_t* out_labels);
/* kmeans_plus_plus_init (exact for n <= 50 000, sampled above)             src/kmeans.rs:154-310 */
int vidx_kmeans_pp_init(int device, const float* data, uint64_t n, uint64_t dim, uint64_t k, uint64_t seed,
                        float* out_centroids);

/* Where the k-means calls of this host thread spent their wall time since the previous call of this function:
 * out[0] = seconds in the serial random stream the reference's semantics keep on the host (a Fisher-Yates shuffle of all n
 * indices per mini-batch iteration, src/kmeans.rs:722-726; a sequential prefix sum per k-means++ draw, :285-287),
 * out[1] = seconds blocked on the device. */
int vidx_kmeans_last_profile(double* out /* 2 */);

/* calculate_num_clusters / calculate_max_iterations                        src/utils.rs:9-26
 * (= suggest_nlist, bindings/python/src/lib.rs:308-315) */
uint64_t vidx_calculate_num_clusters(uint64_t num_vectors);
uint64_t vidx_calculate_max_iterations(uint64_t num_vectors);

/* The random stream every build decision is drawn from (csrc/rng.hpp: rand 0.8.5 StdRng::seed_from_u64 = ChaCha12 behind
 * rand_core's BlockRng; call sites src/kmeans.rs:31,80,170,240,591), host only, so that it can be pinned against known-answer
 * vectors without a GPU.  After `skip_u32` next_u32 draws: kind 0 = n x next_u32, 1 = n x next_u64, 2 = n x gen_range(0..arg)
 * on usize, 3 = (0..arg).shuffle (n == arg), 4 = (0..arg).choose_multiple(n), 5 = the first n entries of (0..arg).shuffle the
 * way the mini-batch loop takes them (every draw made, no n-element array permuted) followed by one next_u32 in out[n].
 * vidx_stdrng_weighted: n samples of WeightedIndex::new(weights) (f32). */
int vidx_stdrng_draw(uint64_t seed, uint64_t skip_u32, int kind, uint64_t arg, uint64_t n, uint64_t* out);
int vidx_stdrng_weighted(uint64_t seed, const float* weights, uint64_t nw, uint64_t n, uint64_t* out);

/* ---- persistence: shard files + index.bin (cold-start path) ----------------------- */

/* IvfIndex::save_to + Shard::save_to for every shard        src/ivf_index.rs:274-294, src/shards.rs:68-177
 * A rank of a shard-partitioned index writes the shard files it owns (rank 0 also index.bin): all ranks saving into
 * the same directories produce the files one GPU would.  Range-partitioned handles cannot save (VIDX_ERR_UNSUPPORTED). */
int vidx_save(const vidx_index* idx, const char* index_dir, const char* shards_dir);
/* VectorIndexer::load + shard load into HBM                  src/api.rs:109-112, src/shards.rs:352-425
 * Missing index.bin => VIDX_ERR_NOT_FOUND; undecodable index.bin => VIDX_ERR_OTHER / VIDX_ERR_INVALID_DATA.
 * A shard file that is missing, has a bad header / shard id / dimension, or whose centroid index does not fit the file
 * is SKIPPED as a whole (its lists stay empty), as search_with_paths drops failed shard reads (src/ivf_index.rs:254);
 * vidx_load_warning_count / vidx_load_warning list what was skipped.  The dimension stored in index.bin replaces the
 * handle's (the reference does not compare it with the config, src/api.rs:109-112): re-read vidx_dimension afterwards.
 * On a partitioned handle (vidx_set_partition before the call) only the owned part is read: the shard files the rank
 * owns, or -- range split -- its vector range of every list. */
int vidx_load(vidx_index* idx, const char* index_dir, const char* shards_dir);
uint64_t vidx_load_warning_count(const vidx_index* idx);
const char* vidx_load_warning(const vidx_index* idx, uint64_t i);

/* The index.bin codec on its own, host only (no device needed): the bincode 2.0.1 `standard()` + ndarray-serde framing of
 * IvfIndex { centroids, centroids_to_shard, dimension } that IvfIndex::save_to / load_index_from use
 * (src/ivf_index.rs:36-41, :274-316).  Read: centroids == NULL && centroids_to_shard == NULL only reports nlist / dimension. */
int vidx_index_bin_write(const char* index_dir, const float* centroids /* nlist x dimension */,
                         const uint64_t* centroids_to_shard /* nlist */, uint64_t nlist, uint32_t dimension);
int vidx_index_bin_read(const char* index_dir, uint64_t cap_lists, float* centroids, uint64_t* centroids_to_shard,
                        uint64_t* nlist, uint32_t* dimension);

/* ---- multi-GPU: shard partition + top-k merge -------------------------------------- */

/* This handle is rank `rank` of `world`: it scans only the lists (or list ranges) it owns.  Shards are dealt to
 * ranks by greedy balance on vector count; the unit is the reference's shard, src/ivf_index.rs:104-164.  Centroids
 * stay replicated so every rank computes the same probe lists.
 *   BEFORE build / load: only the owned part ever reaches HBM (per-rank residency ~ 1/world; training still sees the
 *     whole data set, so every rank derives the same centroids and the same partition).  Such a handle cannot be
 *     re-partitioned afterwards (VIDX_ERR_INVALID_INPUT).
 *   AFTER build / load of a whole index: a mask only -- everything stays resident, any (rank, world) may follow. */
int vidx_set_partition(vidx_index* idx, int rank, int world);
/* How the index is split: 0 = auto (default: shards, unless the most loaded rank would exceed the
 * mean by more than 15 % in vectors or 25 % in expected scan work -- the sum of len^2 over its lists:
 * a query probes a list about as often as a vector falls into it -- then ranges), 1 = shards,
 * 2 = ranges (every rank owns the r-th contiguous
 * range of the segments of every list).  vidx_get_partition_kind: what the last vidx_set_partition
 * used (1 shards, 2 ranges). */
int vidx_set_partition_mode(vidx_index* idx, int mode);
int vidx_get_partition_kind(const vidx_index* idx);
int vidx_get_shard_owner(const vidx_index* idx, int world, int32_t* out /* num_shards */);
/* The partition rule on its own (host only, no device needed): shard_sizes[num_shards]
 * vector counts -> owner rank per shard; largest shard first to the least loaded rank,
 * ties to the lower rank / lower shard id. */
int vidx_partition_shards(const uint64_t* shard_sizes, uint64_t num_shards, int world, int32_t* out);
/* The whole split decision on its own (host only): list sizes + the list -> shard map -> owner rank per shard and *kind = 1
 * (shards) or 2 (ranges), by the rule of vidx_set_partition_mode(mode).  Every rank evaluates it on the same numbers. */
int vidx_partition_plan(const uint64_t* list_sizes, const uint64_t* list_shard, uint64_t nlist, uint64_t num_shards, int world,
                        int mode, int32_t* shard_owner /* num_shards, may be NULL */, int* kind);
/* Merge `nruns` per-rank results (each nq x k, ascending, padded with +inf / -1) laid out run-major in
 * DEVICE memory into the global top-k: the reduction behind join_all + concat + sort
 * in src/ivf_index.rs:249-266.  Any k.  Ties resolve to the lower run index. */
int vidx_merge_topk_device(int device, const float* d_D_runs, const int64_t* d_I_runs, uint32_t nruns, uint64_t nq,
                           uint64_t k, float* d_D, int64_t* d_I, void* stream);
/* For hosts that run the exchange themselves (MPI, one process driving several GPUs): the local part of a partitioned
 * search with, per result, the key (probe rank << 32 | global row) -- padded with UINT64_MAX -- and the merge that
 * orders by (distance, key), i.e. exactly as the reference's stable sort over candidates gathered in probe order does
 * (src/ivf_index.rs:249-266): the merged answer equals the single-GPU answer bit for bit, ties included.
 * vidx_search_multi is these two around one NCCL all-gather. */
int vidx_search_local_device(vidx_index* idx, const float* d_xq, uint64_t nq, uint64_t k, uint64_t n_probe, float* d_D,
                             int64_t* d_I, uint64_t* d_keys, void* stream);
int vidx_merge_topk_keyed_device(int device, const float* d_D_runs, const int64_t* d_I_runs, const uint64_t* d_K_runs,
                                 uint32_t nruns, uint64_t nq, uint64_t k, float* d_D, int64_t* d_I, void* stream);
/* The rank grid of vidx_search_multi on its own (host only): which queries rank `rank` of `world` answers when the index is
 * cut into `parts` parts (world % parts == 0).  out[0] = query group (rank / parts), out[1], out[2] = the group's query
 * range [lo, hi), out[3] = queries per group (the row count of a packed run), out[4], out[5] = the range whose probe lists
 * this rank computes.  vidx_merge_topk_grid_device: the merge behind it -- run g * parts + p (per_group x k entries each,
 * run-major) holds part p's results for the queries of group g; output = the whole batch. */
int vidx_grid_plan(uint64_t nq, int world, int parts, int rank, uint64_t* out /* 6 */);
int vidx_merge_topk_grid_device(int device, const float* d_D_runs, const int64_t* d_I_runs, const uint64_t* d_K_runs, uint32_t parts,
                                uint64_t per_group, uint64_t nq, uint64_t k, float* d_D, int64_t* d_I, void* stream);

/* The exchange step inside the library (north_star 4; replaces join_all + concat + sort, src/ivf_index.rs:249-266):
 * one NCCL communicator per handle, one rank per GPU.  Rank 0 obtains an id (vidx_comm_unique_id), the host program
 * hands it to every rank (any channel: MPI, a file, torch.distributed), every rank calls vidx_comm_init -- collectively.
 * The communicator's world must be a multiple of the handle's partition world P (vidx_set_partition) and rank % P the
 * partition rank: the ranks then form a grid of P index parts x world / P query groups (below).  NCCL is bound with
 * dlopen("libnccl.so.2") at the first call. */
#define VIDX_COMM_ID_BYTES 128
int vidx_comm_unique_id(uint8_t* out /* VIDX_COMM_ID_BYTES */);
int vidx_comm_init(vidx_index* idx, int rank, int world, const uint8_t* unique_id);
int vidx_comm_destroy(vidx_index* idx);
const char* vidx_comm_version(const vidx_index* idx); /* "NCCL x.y.z", "" before vidx_comm_init */
/* Collective search: every rank passes the SAME queries and gets the FULL answer.  Coarse quantization is split by
 * query (all-gather of the probe lists), every rank scans what it owns, the per-rank top-k runs are exchanged with ONE
 * packed all-gather (distance | id | (probe rank, global row) key) and merged on the device by (distance, key) -- the
 * result is bit-identical to a single-GPU vidx_search, ties included, for any k.  Host buffers / device buffers.
 * Grid: with a communicator of world = P x G ranks over a P-way partition, rank r holds index part r % P and answers only
 * the queries of group r / P (a contiguous slice of the batch; the host entry point uploads only that slice); P = world
 * is the plain sharded index, P = 1 (an unpartitioned handle on every rank) a replicated index with the batch split by
 * query -- what a small index on many GPUs wants. */
int vidx_search_multi(vidx_index* idx, const float* xq, uint64_t nq, uint64_t k, uint64_t n_probe, float* D, int64_t* I);
int vidx_search_multi_device(vidx_index* idx, const float* d_xq, uint64_t nq, uint64_t k, uint64_t n_probe, float* d_D,
                             int64_t* d_I, void* stream);

/* ---- measurement ------------------------------------------------------------------- */

typedef struct vidx_search_stats {
    double ms_coarse, ms_select, ms_group, ms_scan, ms_merge, ms_total; /* CUDA-event times of the last search */
    uint64_t scan_bytes_algorithmic; /* sum over distinct probed lists of len*(4D+8) + queries + probes + outputs */
    uint64_t scan_bytes_logical;     /* sum over (query,list) pairs of len*(4D+8)  */
    uint64_t scan_flops;             /* 3*D per (query,vector) pair (sub, mul, add) */
    uint64_t coarse_flops;           /* 3*D*nq*nlist */
    uint64_t n_pairs;                /* (query, segment) pairs scanned */
    uint64_t n_dense_items, n_sparse_items;
    uint64_t kernel_launches;        /* kernels launched by the last search */
    double ms_scan_tc;               /* the tensor-core scan kernel alone (part of ms_scan) */
    uint64_t n_tc_items, n_tc_survivors, n_tc_overflow; /* work items, candidates re-checked exactly, queries redone exactly */
    uint64_t tc_mma_flops;           /* 2*D per (query, vector) pair issued to the tensor cores */
    uint64_t n_tc_submin_slots;       /* sub-tile minima reserved for the bounds pass (0 = the seeded flavour ran) */
} vidx_search_stats;
/* Enable per-stage CUDA-event timing (adds stream synchronisation at the end of a
 * search); stats describe the last completed search on this handle. */
int vidx_set_profiling(vidx_index* idx, int enabled);
/* Scan algorithm: 0 (default) = tcgen05 FP16 filter + exact re-check whenever the shape allows (k <= 32, D <= 2048, finite data;
 * up to D = 512 the query tile stays in shared memory, beyond it is streamed through the ring with the list tiles).
 * The filter runs a bounds pass (minima only) and a main pass: when a query visits many 128-vector tiles (tensor-bound) the
 * bounds pass covers the heads of its nearest lists and the main pass keeps tightening the bounds; when it visits at most
 * 2048 tiles (the HBM-bound regime) the bounds pass covers everything and the main pass only collects.
 * 1 = exact FP32 kernels only; 2 / 3 = force the first / the second flavour of the filter (the second whenever its minima fit
 * in 8 GB).  Results are bit-identical in every mode. */
int vidx_set_scan_mode(vidx_index* idx, int mode);
/* Coarse quantization (ivf_index.rs:205-220): 0 = auto (the tensor-core filter + exact re-check when n_probe <= 32, the table has
 * >= 512 lists and the batch >= 8M (query, centroid) pairs -- measured faster from there, 12x at nlist = 65 536; else the exact
 * FP32 kernels), 1 = exact kernels only, 2 = the filter whenever it applies (n_probe <= 32).  Probe lists and distances are
 * identical in every mode. */
int vidx_set_coarse_mode(vidx_index* idx, int mode);
int vidx_get_search_stats(vidx_index* idx, vidx_search_stats* out);
/* Total kernels launched by this library in this process (bench.py's gpu_launches). */
uint64_t vidx_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VIDX_B200_H */

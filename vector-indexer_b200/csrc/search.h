// search.h -- declarations shared by search_kernels.cu and the index host code.
#pragma once
#include "common.cuh"

namespace vidx {

// One unit of scan work: a tile of the queries that probe segment `seg`.
struct ScanItem {
    uint32_t seg;
    uint32_t qstart;  // offset into the segment's query list
    uint32_t nq;      // <= 64 (dense) or <= 8 (sparse)
    uint32_t pad;
};

void exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, size_t n, uint32_t* d_tmp, cudaStream_t st);
size_t exclusive_scan_tmp_entries(size_t n);

void launch_fill_u32(uint32_t* p, uint32_t v, size_t n, cudaStream_t st);
void launch_pad_rows(const float* in, float* out, uint64_t nrows, int D, int Dp, cudaStream_t st);
void launch_interleave(const float* src, int D, int Dq, const uint32_t* row_src, size_t nrows, float4* dst, cudaStream_t st);
void launch_coarse_dist(const float4* cents, int ngroups, int Dq, const float4* xq4, uint32_t nq, float* out, uint32_t ldo,
                        cudaStream_t st);
uint32_t select_kcap(uint32_t k);
void launch_select_topk(const float* vals, const uint64_t* row_off, const uint32_t* row_len, uint64_t ld, uint32_t n_fixed,
                        uint64_t nrows, uint32_t k, uint32_t* out_pos, float* out_val, cudaStream_t st);
// the same for k <= 32 and short rows (one warp per row)
void launch_select_small(const float* vals, const uint64_t* row_off, const uint32_t* row_len, uint64_t ld, uint32_t n_fixed,
                        uint64_t nrows, uint32_t k, uint32_t* out_pos, float* out_val, cudaStream_t st);
void launch_group_count(const uint32_t* probes, size_t npairs, uint32_t nprobe, const uint2* list_seg,
                        const uint32_t* only_flag, uint32_t* pair_ns, uint32_t* seg_cnt, cudaStream_t st);
void launch_group_fill(const uint32_t* probes, size_t npairs, uint32_t nprobe, const uint2* list_seg,
                       const uint32_t* slot_off, const uint32_t* seg_qoff, uint32_t* seg_cur, uint2* seg_qlist,
                       uint32_t* slot_seg, const uint32_t* only_flag, uint32_t* slot_rank, cudaStream_t st);
void launch_group_items(const uint32_t* seg_cnt, uint32_t nseg, uint32_t sparse_max, ScanItem* dense, ScanItem* sparse,
                        uint32_t* counters, cudaStream_t st);
bool sparse_supported(int Dq);
void launch_scan(bool alldist, const float4* vecs, int Dq, const float4* xq4, const SegDesc* segs, const uint32_t* seg_qoff,
                 const uint2* seg_qlist, const ScanItem* dense, const ScanItem* sparse, const uint32_t* counters,
                 uint32_t* work_counters, uint32_t k, float* cand_d, uint32_t* cand_r, float* alld, bool use_sparse,
                 cudaStream_t st);
void launch_merge_slots(const float* cand_d, const uint32_t* cand_r, const uint32_t* slot_off, uint32_t nq, uint32_t nprobe,
                        uint32_t k, uint32_t kout, const uint64_t* row_ext, float* D, int64_t* I, uint32_t* out_rows,
                        cudaStream_t st);
void launch_pad_output(float* D, int64_t* I, uint32_t* rows, unsigned long long* keys, uint64_t nq, uint32_t k, uint32_t kout,
                       cudaStream_t st);
void launch_alldist_rows(const uint32_t* slot_off, uint32_t nprobe, uint64_t nq, uint64_t* row_off, uint32_t* row_len,
                         cudaStream_t st);
void launch_alldist_finish(const uint32_t* sel_pos, const float* sel_val, const uint32_t* slot_off, const uint32_t* slot_seg,
                           const SegDesc* segs, uint32_t nprobe, uint64_t nq, uint32_t k, uint32_t kout,
                           const uint64_t* row_ext, float* D, int64_t* I, uint32_t* out_rows, const uint32_t* slot_rank,
                           const uint32_t* list_rowdelta, unsigned long long* out_keys, cudaStream_t st);
void launch_gather_vectors(const float* vecs, int Dq, int D, const uint32_t* rows, size_t nres, float* out, cudaStream_t st);
// runs: D at Dr + r * sD, I at Ir + r * sI, optional keys at Kr + r * sK (strides in elements); any k; see merge_runs_kernel.
// per_group > 0: queries come in groups of per_group, group g's runs are runs g * nruns .. and hold that group's rows only.
void launch_merge_runs(const float* Dr, size_t sD, const int64_t* Ir, size_t sI, const unsigned long long* Kr, size_t sK,
                       uint32_t nruns, uint64_t nq, uint32_t k, float* D, int64_t* I, cudaStream_t st, uint64_t per_group = 0);

}  // namespace vidx

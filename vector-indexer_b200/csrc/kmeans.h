// kmeans.h -- declarations shared by kmeans_kernels.cu and the k-means host drivers.
#pragma once
#include "common.cuh"

namespace vidx {

// A batch of (points x centroids) distance work: points pt_off..pt_off+npts of the point
// entry list against centroids c_off..c_off+ncents of the centroid index list.
struct PairItem {
    uint32_t pt_off, npts, c_off, ncents;
};

void launch_pairs(int mode, const float* data, int D, const float* cents, const PairItem* d_items,
                  const uint32_t* d_item_tile_off, int nitems, uint32_t total_tiles, const uint2* pt_entries,
                  const uint32_t* cent_idx, float* out, uint64_t ld, unsigned long long* best, cudaStream_t st);
uint32_t pairs_point_tile();
void launch_top3(const float* dist, uint64_t ld, uint32_t npts, uint32_t meta_k, uint32_t top, uint32_t* out, cudaStream_t st);
void launch_meta_count(const uint32_t* top3, uint32_t npts, uint32_t* cnt, cudaStream_t st);
void launch_meta_fill(const uint32_t* top3, uint32_t npts, uint32_t pt_base, const uint32_t* off, uint32_t* cur,
                      uint2* entries, cudaStream_t st);
void launch_keys_to_labels(const unsigned long long* best, uint32_t npts, const uint32_t* top3, const uint32_t* m2c_off,
                           const uint32_t* m2c_list, uint32_t* labels, cudaStream_t st);
void launch_fill_keys(unsigned long long* p, size_t n, cudaStream_t st);
void launch_min_dist(const float* data, int D, uint32_t m, const float* latest, float* min_d, cudaStream_t st);
void launch_copy_rows(const float* src, const uint32_t* src_idx, float* dst, const uint32_t* dst_idx, uint32_t nrows, int D,
                      cudaStream_t st);
void launch_cluster_mean(const float* data, int D, const uint32_t* member_off, const uint32_t* members,
                         const uint32_t* cluster_ids, const float* eta, uint32_t nclusters, int mode, float* out,
                         cudaStream_t st);
void launch_centroid_delta(const float* curr, const float* prev, uint32_t k, int D, float* local, cudaStream_t st);

}  // namespace vidx

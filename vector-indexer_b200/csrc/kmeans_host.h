// kmeans_host.h -- device k-means drivers (mirror of the functions src/kmeans.rs exports).
#pragma once
#include <vector>

#include "kmeans.h"

namespace vidx {

// Wall-time split of the k-means calls made on this host thread since the last reset (vidx_kmeans_last_profile).
struct KmProfile {
    double host_rng_s = 0.0;     // the reference's serial random stream: shuffles over all n, weighted draws with their prefix sums
    double device_wait_s = 0.0;  // blocked on the stream (kernels + copies)
};
extern thread_local KmProfile t_km_profile;

// Two-level centroid hierarchy (kmeans.rs:584-648).
struct Hierarchy {
    uint32_t meta_k = 0;
    std::vector<uint32_t> c2m;       // centroid -> meta
    std::vector<uint32_t> m2c_off;   // meta -> [centroids], ascending
    std::vector<uint32_t> m2c_list;
};

// K-means over an n x D row-major fp32 matrix that already sits in device memory.
class DeviceKMeans {
  public:
    struct Impl;
    DeviceKMeans(const float* d_data, uint64_t n, int D, cudaStream_t st);
    ~DeviceKMeans();
    DeviceKMeans(const DeviceKMeans&) = delete;
    DeviceKMeans& operator=(const DeviceKMeans&) = delete;

    // kmeans_plus_plus_init (kmeans.rs:154-310) -> d_cents[k][D]
    void pp_init(uint32_t k, uint64_t seed, float* d_cents);
    // run_kmeans_mini_batch (kmeans.rs:64-150); returns iterations run
    uint64_t mini_batch(uint32_t k, uint64_t max_iters, float tol, uint64_t seed, float* d_cents, uint32_t* d_labels);
    // run_kmeans_parallel (kmeans.rs:15-60)
    uint64_t lloyd(uint32_t k, uint64_t max_iters, float tol, uint64_t seed, float* d_cents, uint32_t* d_labels);
    // assign_points_simd_parallel (kmeans.rs:445-459) over the whole data set
    void assign(const float* d_cents, uint32_t k, uint64_t seed, uint32_t* d_labels);

    void assign_brute(const float* d_pts, uint64_t npts, const float* d_cents, uint32_t k, uint32_t* d_labels);
    void assign_hierarchical(const float* d_pts, uint64_t npts, const float* d_cents, uint32_t k, uint64_t seed,
                             uint32_t* d_labels);
    void build_hierarchy(const float* d_cents, uint32_t k, uint64_t hseed, Hierarchy& h);

  private:
    const float* d_data_;
    uint64_t n_;
    int D_;
    cudaStream_t st_;
    Impl* impl_;
};

}  // namespace vidx

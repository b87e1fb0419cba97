// kmeans_host.cu -- host drivers of k-means (src/kmeans.rs): control flow, the random
// stream and the small sequential pieces stay on the host; every distance, argmin and
// centroid update runs on the device.
#include <chrono>

#include "kmeans_host.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

#include "rng.hpp"

namespace vidx {

thread_local KmProfile t_km_profile;

namespace {

template <class T>
void h2d(T* dst, const T* src, size_t n, cudaStream_t st) {
    if (n) VIDX_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
}
// Where a k-means call spends its wall time: the serial host parts the reference's semantics force (the rand stream:
// Fisher-Yates over all n indices per iteration, one sequential 50 000-term prefix sum per k-means++ draw) against
// waiting for the device.  Per host thread; read through vidx_kmeans_last_profile.
struct Tick {
    double& acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit Tick(double& a) : acc(a) {}
    ~Tick() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
#define VIDX_SYNC(st)                                  \
    do {                                               \
        Tick _t(t_km_profile.device_wait_s);           \
        VIDX_CUDA(cudaStreamSynchronize(st));          \
    } while (0)

template <class T>
void d2h(T* dst, const T* src, size_t n, cudaStream_t st) {
    if (n) VIDX_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, st));
    VIDX_SYNC(st);
}

// Single-item pair launch (all `npts` points against all `k` centroids).
struct FlatItem {
    DevBuf items, tile_off;
    void launch(int mode, const float* d_pts, uint32_t npts, int D, const float* d_cents, uint32_t k, float* out,
                uint64_t ld, unsigned long long* best, cudaStream_t st) {
        if (!npts || !k) return;
        PairItem it{0, npts, 0, k};
        uint32_t tiles = (uint32_t)ceil_div(npts, pairs_point_tile());
        uint32_t off[2] = {0, tiles};
        items.reserve(sizeof(it));
        tile_off.reserve(sizeof(off));
        h2d(items.as<PairItem>(), &it, 1, st);
        h2d(tile_off.as<uint32_t>(), off, 2, st);
        // (pageable-source cudaMemcpyAsync stages the bytes before returning, so the stack
        // temporaries may go out of scope)
        launch_pairs(mode, d_pts, D, d_cents, items.as<PairItem>(), tile_off.as<uint32_t>(), 1, tiles, nullptr, nullptr, out,
                     ld, best, st);
    }
};

}  // namespace

struct DeviceKMeans::Impl {
    FlatItem flat;
    DevBuf keys, labels32, distm, top3, cnt, cur, off, entries, items, tile_off, m2c_off, m2c_list, meta, members, member_off,
        cluster_ids, eta, local, batch, idx_a, idx_b, prev, mind, newc;
    PinnedBuf pin;
};

DeviceKMeans::DeviceKMeans(const float* d_data, uint64_t n, int D, cudaStream_t st)
    : d_data_(d_data), n_(n), D_(D), st_(st), impl_(new Impl) {}
DeviceKMeans::~DeviceKMeans() { delete impl_; }

// ---- brute-force argmin of every point over all k centroids (kmeans.rs:462-470) ------
void DeviceKMeans::assign_brute(const float* d_pts, uint64_t npts, const float* d_cents, uint32_t k, uint32_t* d_labels) {
    Impl& m = *impl_;
    const uint64_t chunk = 1ull << 24;
    for (uint64_t p0 = 0; p0 < npts; p0 += chunk) {
        uint32_t np = (uint32_t)std::min(chunk, npts - p0);
        m.keys.reserve((size_t)np * 8);
        launch_fill_keys(m.keys.as<unsigned long long>(), np, st_);
        m.flat.launch(1, d_pts + p0 * D_, np, D_, d_cents, k, nullptr, 0, m.keys.as<unsigned long long>(), st_);
        launch_keys_to_labels(m.keys.as<unsigned long long>(), np, nullptr, nullptr, nullptr, d_labels + p0, st_);
    }
}

// ---- centroid hierarchy (kmeans.rs:584-648) -------------------------------------------
void DeviceKMeans::build_hierarchy(const float* d_cents, uint32_t k, uint64_t hseed, Hierarchy& h) {
    Impl& m = *impl_;
    uint32_t meta_k = (uint32_t)std::min<uint64_t>(std::max<uint64_t>((uint64_t)std::sqrt((float)k), 2), k / 2);
    h.meta_k = meta_k;
    ChaCha12Rng rng(hseed);
    std::vector<uint32_t> chosen = rng.choose_multiple(k, meta_k);
    m.meta.reserve((size_t)meta_k * D_ * 4);
    m.idx_a.reserve((size_t)std::max<uint32_t>(k, meta_k) * 4);
    VIDX_CUDA(cudaMemsetAsync(m.meta.p, 0, (size_t)meta_k * D_ * 4, st_));
    h2d(m.idx_a.as<uint32_t>(), chosen.data(), chosen.size(), st_);
    launch_copy_rows(d_cents, m.idx_a.as<uint32_t>(), m.meta.as<float>(), nullptr, (uint32_t)chosen.size(), D_, st_);
    m.labels32.reserve((size_t)std::max<uint64_t>(k, 1) * 4);
    h.c2m.assign(k, 0);
    std::vector<uint32_t> moff(meta_k + 1), mem(k);
    m.member_off.reserve((size_t)(meta_k + 1) * 4);
    m.members.reserve((size_t)k * 4);
    for (int iter = 0; iter < 5; iter++) {
        assign_brute(d_cents, k, m.meta.as<float>(), meta_k, m.labels32.as<uint32_t>());
        d2h(h.c2m.data(), m.labels32.as<uint32_t>(), k, st_);
        std::fill(moff.begin(), moff.end(), 0u);
        for (uint32_t c = 0; c < k; c++) moff[h.c2m[c] + 1]++;
        for (uint32_t t = 0; t < meta_k; t++) moff[t + 1] += moff[t];
        std::vector<uint32_t> cur(moff.begin(), moff.end() - 1);
        for (uint32_t c = 0; c < k; c++) mem[cur[h.c2m[c]]++] = c;
        h2d(m.member_off.as<uint32_t>(), moff.data(), moff.size(), st_);
        h2d(m.members.as<uint32_t>(), mem.data(), mem.size(), st_);
        launch_cluster_mean(d_cents, D_, m.member_off.as<uint32_t>(), m.members.as<uint32_t>(), nullptr, nullptr, meta_k, 1,
                            m.meta.as<float>(), st_);
        VIDX_SYNC(st_);
    }
    // meta -> centroid lists, centroids ascending inside each meta (kmeans.rs:518-521);
    // moff/mem of the last iteration are exactly that.
    h.m2c_off = moff;
    h.m2c_list = mem;
    m.m2c_off.reserve(moff.size() * 4);
    m.m2c_list.reserve(std::max<size_t>(mem.size(), 1) * 4);
    h2d(m.m2c_off.as<uint32_t>(), moff.data(), moff.size(), st_);
    h2d(m.m2c_list.as<uint32_t>(), mem.data(), mem.size(), st_);
    VIDX_SYNC(st_);
}

// ---- hierarchical assignment (kmeans.rs:474-581) --------------------------------------
void DeviceKMeans::assign_hierarchical(const float* d_pts, uint64_t npts, const float* d_cents, uint32_t k, uint64_t seed,
                                       uint32_t* d_labels) {
    Impl& m = *impl_;
    Hierarchy h;
    build_hierarchy(d_cents, k, seed * 17ull + 42ull, h);
    uint32_t meta_k = h.meta_k;
    uint32_t top = std::min<uint32_t>(3, meta_k);
    const uint64_t chunk = 1ull << 20;
    std::vector<uint32_t> cnt(meta_k), off(meta_k + 1), tile_off;
    std::vector<PairItem> items;
    for (uint64_t p0 = 0; p0 < npts; p0 += chunk) {
        uint32_t np = (uint32_t)std::min(chunk, npts - p0);
        const float* pts = d_pts + p0 * D_;
        // stage 1: distances to the meta-centroids, stable top-3
        m.distm.reserve((size_t)np * meta_k * 4);
        m.flat.launch(0, pts, np, D_, m.meta.as<float>(), meta_k, m.distm.as<float>(), meta_k, nullptr, st_);
        m.top3.reserve((size_t)np * 3 * 4);
        launch_top3(m.distm.as<float>(), meta_k, np, meta_k, top, m.top3.as<uint32_t>(), st_);
        // group (point, rank) by meta
        m.cnt.reserve((size_t)meta_k * 4);
        m.cur.reserve((size_t)meta_k * 4);
        m.off.reserve((size_t)(meta_k + 1) * 4);
        VIDX_CUDA(cudaMemsetAsync(m.cnt.p, 0, (size_t)meta_k * 4, st_));
        VIDX_CUDA(cudaMemsetAsync(m.cur.p, 0, (size_t)meta_k * 4, st_));
        launch_meta_count(m.top3.as<uint32_t>(), np, m.cnt.as<uint32_t>(), st_);
        d2h(cnt.data(), m.cnt.as<uint32_t>(), meta_k, st_);
        off[0] = 0;
        for (uint32_t t = 0; t < meta_k; t++) off[t + 1] = off[t] + cnt[t];
        h2d(m.off.as<uint32_t>(), off.data(), off.size(), st_);
        m.entries.reserve((size_t)std::max<uint32_t>(off[meta_k], 1) * 8);
        launch_meta_fill(m.top3.as<uint32_t>(), np, 0, m.off.as<uint32_t>(), m.cur.as<uint32_t>(), m.entries.as<uint2>(), st_);
        // stage 2: per meta, its points against its centroids; argmin in candidate order
        items.clear();
        tile_off.assign(1, 0u);
        for (uint32_t t = 0; t < meta_k; t++) {
            uint32_t nc = h.m2c_off[t + 1] - h.m2c_off[t];
            if (!cnt[t] || !nc) continue;
            items.push_back(PairItem{off[t], cnt[t], h.m2c_off[t], nc});
            tile_off.push_back(tile_off.back() + (uint32_t)ceil_div(cnt[t], pairs_point_tile()));
        }
        m.keys.reserve((size_t)np * 8);
        launch_fill_keys(m.keys.as<unsigned long long>(), np, st_);
        if (!items.empty()) {
            m.items.reserve(items.size() * sizeof(PairItem));
            m.tile_off.reserve(tile_off.size() * 4);
            h2d(m.items.as<PairItem>(), items.data(), items.size(), st_);
            h2d(m.tile_off.as<uint32_t>(), tile_off.data(), tile_off.size(), st_);
            launch_pairs(1, pts, D_, d_cents, m.items.as<PairItem>(), m.tile_off.as<uint32_t>(), (int)items.size(),
                         tile_off.back(), m.entries.as<uint2>(), m.m2c_list.as<uint32_t>(), nullptr, 0,
                         m.keys.as<unsigned long long>(), st_);
        }
        launch_keys_to_labels(m.keys.as<unsigned long long>(), np, m.top3.as<uint32_t>(), m.m2c_off.as<uint32_t>(),
                              m.m2c_list.as<uint32_t>(), d_labels + p0, st_);
        VIDX_SYNC(st_);  // host vectors reused next chunk
    }
}

// kmeans.rs:445-459
void DeviceKMeans::assign(const float* d_cents, uint32_t k, uint64_t seed, uint32_t* d_labels) {
    if (k > 100) assign_hierarchical(d_data_, n_, d_cents, k, seed, d_labels);
    else assign_brute(d_data_, n_, d_cents, k, d_labels);
}

// ---- k-means++ (kmeans.rs:154-310) -----------------------------------------------------
void DeviceKMeans::pp_init(uint32_t k, uint64_t seed, float* d_cents) {
    Impl& m = *impl_;
    ChaCha12Rng rng(seed);
    const uint64_t sample_threshold = 50000;
    const bool sampled = n_ > sample_threshold;
    const uint64_t actual_k = std::min<uint64_t>(k, n_);
    VIDX_CUDA(cudaMemsetAsync(d_cents, 0, (size_t)k * D_ * 4, st_));
    auto copy_data_row = [&](uint64_t src, uint32_t dst) {
        VIDX_CUDA(cudaMemcpyAsync(d_cents + (size_t)dst * D_, d_data_ + src * D_, (size_t)D_ * 4, cudaMemcpyDeviceToDevice, st_));
    };
    auto copy_cent_row = [&](uint32_t src, uint32_t dst) {
        VIDX_CUDA(cudaMemcpyAsync(d_cents + (size_t)dst * D_, d_cents + (size_t)src * D_, (size_t)D_ * 4, cudaMemcpyDeviceToDevice,
                                  st_));
    };
    uint64_t first = rng.below_u64(n_);
    copy_data_row(first, 0);
    std::vector<uint32_t> sample_idx;
    uint64_t mcount = n_;
    if (sampled) {
        sample_idx.resize(n_);
        std::iota(sample_idx.begin(), sample_idx.end(), 0u);
        {
            Tick t(t_km_profile.host_rng_s);
            rng.shuffle(sample_idx.data(), sample_idx.size());
        }
        sample_idx.resize(sample_threshold);
        mcount = sample_threshold;
    }
    m.mind.reserve(mcount * 4);
    m.pin.reserve(mcount * 4);
    float* h_min = m.pin.as<float>();
    {
        std::vector<float> inf(mcount, std::numeric_limits<float>::infinity());
        h2d(m.mind.as<float>(), inf.data(), mcount, st_);
        VIDX_SYNC(st_);
    }
    std::vector<float> cum;
    for (uint64_t i = 1; i < actual_k; i++) {
        // NB kmeans.rs:268/:431-436: rows 0..m of the data, also in the sampled variant.
        launch_min_dist(d_data_, D_, (uint32_t)mcount, d_cents + (i - 1) * D_, m.mind.as<float>(), st_);
        d2h(h_min, m.mind.as<float>(), mcount, st_);
        Tick tk(t_km_profile.host_rng_s);  // the draw: one sequential f32 chain over all weights (WeightedIndex, kmeans.rs:270-287)
        const float total = ChaCha12Rng::square_prefix(h_min, mcount, cum);
        if (total == 0.0f) {
            copy_cent_row((uint32_t)rng.below_u64(i), (uint32_t)i);
        } else {
            size_t s = rng.pick_from_prefix(cum, total);
            copy_data_row(sampled ? sample_idx[s] : s, (uint32_t)i);
        }
    }
    for (uint64_t i = actual_k; i < k; i++) copy_cent_row((uint32_t)rng.below_u64(actual_k), (uint32_t)i);
    VIDX_SYNC(st_);
}

// kmeans.rs:313-331: clusters with count 0 take a fresh random data row, ascending c.
static void reseed_empty(DeviceKMeans::Impl& m, const float* d_data, uint64_t n, int D, float* d_cents,
                         const std::vector<uint64_t>& counts, ChaCha12Rng& rng, cudaStream_t st) {
    std::vector<uint32_t> cs, rs;
    for (size_t c = 0; c < counts.size(); c++)
        if (counts[c] == 0) {
            cs.push_back((uint32_t)c);
            rs.push_back((uint32_t)rng.below_u64(n));
        }
    if (cs.empty()) return;
    m.idx_a.reserve(cs.size() * 4);
    m.idx_b.reserve(rs.size() * 4);
    h2d(m.idx_a.as<uint32_t>(), cs.data(), cs.size(), st);
    h2d(m.idx_b.as<uint32_t>(), rs.data(), rs.size(), st);
    launch_copy_rows(d_data, m.idx_b.as<uint32_t>(), d_cents, m.idx_a.as<uint32_t>(), (uint32_t)cs.size(), D, st);
    VIDX_SYNC(st);
}

// kmeans.rs:334-351; cross-centroid order fixed to ascending c (the reference's rayon sum
// has no defined order).
static float centroid_delta(DeviceKMeans::Impl& m, const float* d_curr, const float* d_prev, uint32_t k, int D,
                            cudaStream_t st) {
    m.local.reserve((size_t)k * 4);
    launch_centroid_delta(d_curr, d_prev, k, D, m.local.as<float>(), st);
    std::vector<float> local(k);
    d2h(local.data(), m.local.as<float>(), k, st);
    float total = 0.0f;
    for (uint32_t c = 0; c < k; c++) total += local[c];
    return std::sqrt(total / (float)((uint64_t)k * (uint64_t)D));
}

// ---- mini-batch k-means (kmeans.rs:64-150) ---------------------------------------------
uint64_t DeviceKMeans::mini_batch(uint32_t k, uint64_t max_iters, float tol, uint64_t seed, float* d_cents,
                                  uint32_t* d_labels) {
    Impl& m = *impl_;
    if (tol < 0) tol = 1e-4f;
    ChaCha12Rng rng(seed);
    const uint64_t batch_size = std::min<uint64_t>(256, std::max<uint64_t>(10, (uint64_t)std::sqrt((float)n_)));
    pp_init(k, seed, d_cents);
    std::vector<uint64_t> counts(k, 0);
    m.prev.reserve((size_t)k * D_ * 4);
    VIDX_CUDA(cudaMemcpyAsync(m.prev.p, d_cents, (size_t)k * D_ * 4, cudaMemcpyDeviceToDevice, st_));
    std::vector<uint32_t> idx, shuffle_scratch;
    const uint32_t b = (uint32_t)std::min<uint64_t>(batch_size, n_);
    m.batch.reserve((size_t)b * D_ * 4);
    m.labels32.reserve((size_t)std::max<uint64_t>(b, 1) * 4);
    std::vector<uint32_t> bl(b), cluster_ids, member_off, members;
    std::vector<float> eta;
    uint64_t it = 0;
    while (it < max_iters) {
        // sample_batch (kmeans.rs:722-726): shuffle all n indices, take the first b
        {
            Tick t(t_km_profile.host_rng_s);  // sample_batch: every draw of a Fisher-Yates over ALL n indices (kmeans.rs:722-726)
            rng.shuffle_head((uint32_t)n_, b, idx, shuffle_scratch);
        }
        m.idx_a.reserve((size_t)b * 4);
        h2d(m.idx_a.as<uint32_t>(), idx.data(), b, st_);
        launch_copy_rows(d_data_, m.idx_a.as<uint32_t>(), m.batch.as<float>(), nullptr, b, D_, st_);
        // assignment of the batch: always brute force over all k (kmeans.rs:103-110)
        assign_brute(m.batch.as<float>(), b, d_cents, k, m.labels32.as<uint32_t>());
        d2h(bl.data(), m.labels32.as<uint32_t>(), b, st_);
        // group by cluster, members in batch order (kmeans.rs:739-742)
        std::vector<std::vector<uint32_t>> pts(k);
        for (uint32_t t = 0; t < b; t++) pts[bl[t]].push_back(idx[t]);
        cluster_ids.clear(); member_off.assign(1, 0u); members.clear(); eta.clear();
        for (uint32_t c = 0; c < k; c++) {
            if (pts[c].empty()) continue;
            uint64_t nc = counts[c] + 1;  // +1 per batch that touched the cluster (kmeans.rs:757)
            counts[c] = nc;
            cluster_ids.push_back(c);
            eta.push_back(1.0f / (float)nc);
            members.insert(members.end(), pts[c].begin(), pts[c].end());
            member_off.push_back((uint32_t)members.size());
        }
        m.cluster_ids.reserve(cluster_ids.size() * 4);
        m.member_off.reserve(member_off.size() * 4);
        m.members.reserve(std::max<size_t>(members.size(), 1) * 4);
        m.eta.reserve(eta.size() * 4);
        h2d(m.cluster_ids.as<uint32_t>(), cluster_ids.data(), cluster_ids.size(), st_);
        h2d(m.member_off.as<uint32_t>(), member_off.data(), member_off.size(), st_);
        h2d(m.members.as<uint32_t>(), members.data(), members.size(), st_);
        h2d(m.eta.as<float>(), eta.data(), eta.size(), st_);
        launch_cluster_mean(d_data_, D_, m.member_off.as<uint32_t>(), m.members.as<uint32_t>(), m.cluster_ids.as<uint32_t>(),
                            m.eta.as<float>(), (uint32_t)cluster_ids.size(), 2, d_cents, st_);
        VIDX_SYNC(st_);
        reseed_empty(m, d_data_, n_, D_, d_cents, counts, rng, st_);
        float delta = centroid_delta(m, d_cents, m.prev.as<float>(), k, D_, st_);
        VIDX_CUDA(cudaMemcpyAsync(m.prev.p, d_cents, (size_t)k * D_ * 4, cudaMemcpyDeviceToDevice, st_));
        it++;
        if (delta < tol) break;
    }
    assign(d_cents, k, seed, d_labels);
    VIDX_SYNC(st_);
    return it;
}

// ---- full-batch Lloyd (kmeans.rs:15-60) -------------------------------------------------
uint64_t DeviceKMeans::lloyd(uint32_t k, uint64_t max_iters, float tol, uint64_t seed, float* d_cents, uint32_t* d_labels) {
    Impl& m = *impl_;
    if (tol < 0) tol = 1e-4f;
    ChaCha12Rng rng(seed);
    pp_init(k, seed, d_cents);
    VIDX_CUDA(cudaMemsetAsync(d_labels, 0, n_ * 4, st_));
    m.newc.reserve((size_t)k * D_ * 4);
    std::vector<uint32_t> labels(n_), member_off(k + 1), members(n_);
    std::vector<uint64_t> counts(k);
    m.member_off.reserve((size_t)(k + 1) * 4);
    m.members.reserve(std::max<uint64_t>(n_, 1) * 4);
    uint64_t it = 0;
    while (it < max_iters) {
        assign(d_cents, k, seed, d_labels);
        d2h(labels.data(), d_labels, n_, st_);
        // members of each cluster in ascending point index (kmeans.rs:681-684)
        std::fill(member_off.begin(), member_off.end(), 0u);
        for (uint64_t i = 0; i < n_; i++) member_off[labels[i] + 1]++;
        for (uint32_t c = 0; c < k; c++) {
            counts[c] = member_off[c + 1];
            member_off[c + 1] += member_off[c];
        }
        std::vector<uint32_t> cur(member_off.begin(), member_off.end() - 1);
        for (uint64_t i = 0; i < n_; i++) members[cur[labels[i]]++] = (uint32_t)i;
        h2d(m.member_off.as<uint32_t>(), member_off.data(), member_off.size(), st_);
        h2d(m.members.as<uint32_t>(), members.data(), members.size(), st_);
        launch_cluster_mean(d_data_, D_, m.member_off.as<uint32_t>(), m.members.as<uint32_t>(), nullptr, nullptr, k, 0,
                            m.newc.as<float>(), st_);
        VIDX_SYNC(st_);
        reseed_empty(m, d_data_, n_, D_, m.newc.as<float>(), counts, rng, st_);
        float delta = centroid_delta(m, m.newc.as<float>(), d_cents, k, D_, st_);
        VIDX_CUDA(cudaMemcpyAsync(d_cents, m.newc.p, (size_t)k * D_ * 4, cudaMemcpyDeviceToDevice, st_));
        it++;
        if (delta < tol) break;
    }
    VIDX_SYNC(st_);
    return it;
}

}  // namespace vidx

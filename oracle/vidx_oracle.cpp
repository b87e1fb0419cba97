// vidx_oracle.cpp -- CPU restatement of the reference's IVF search + k-means path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (vector-indexer_b200/) may
// include, link, import or execute this file; only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs use it, as the checker or
// as the timed CPU baseline -- never as the thing shipped.
//
// PARITY STATUS: "parity unpinned" for every RNG-derived quantity.  The reference
// (NirajNair/vector-indexer, pure Rust) cannot be compiled here (no cargo/rustc,
// crates not vendored, no network) and its test-suite holds no golden vector or
// known-answer value.  What IS pinned:
//   * the ChaCha block function, checked at 20 rounds against OpenSSL (via the
//     `cryptography` package) in tests/test_oracle_rng.py; the 12-round StdRng is
//     the same code with a different round count;
//   * the deterministic fixtures of the reference tests (ramp vectors
//     tests/test_utils/mod.rs:10-16, make_records tests/api_tests.rs:12-25, shard
//     byte layouts src/shards.rs:22-51) and all the reference's test *properties*.
// Third-party behaviour restated from the published crates pinned in Cargo.lock
// (sources are not on disk): rand 0.8.5, rand_chacha 0.3.1, rand_core 0.6.4,
// wide 0.7.33 (non-AVX build: f32x8::reduce_add = a.reduce_add() + b.reduce_add(),
// f32x4::reduce_add = sequential sum).
//
// Build: see oracle/Makefile  (-O3 -ffp-contract=off, no fast-math: Rust never
// contracts mul+add into FMA and never reassociates float sums).
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef uint32_t u32;

// ===========================================================================
// rand 0.8.5 / rand_chacha 0.3.1 / rand_core 0.6.4  
// ===========================================================================
static inline u32 rotl32(u32 x, int r) { return (x << r) | (x >> (32 - r)); }

// One ChaCha block ("djb" variant: 64-bit counter in words 12-13, 64-bit stream in
// 14-15).  rounds = 12 for StdRng (rand 0.8.5: StdRng = ChaCha12Rng).
static void chacha_block(const u32 key[8], u64 counter, u64 stream, int rounds, u32 out[16]) {
    u32 in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                  key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                  (u32)counter, (u32)(counter >> 32), (u32)stream, (u32)(stream >> 32)};
    u32 x[16];
    memcpy(x, in, sizeof(x));
#define QR(a, b, c, d)                  \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
    for (int r = 0; r < rounds; r += 2) {
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#undef QR
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

// BlockRng<ChaCha12Core>: 64-word buffer refilled 4 blocks at a time.
struct StdRng {
    u32 key[8];
    u64 counter;
    u32 buf[64];
    int index;

    // rand_core 0.6.4 SeedableRng::seed_from_u64: PCG32 expansion of the u64.
    explicit StdRng(u64 state) {
        for (int i = 0; i < 8; i++) {
            state = state * 6364136223846793005ULL + 11634580027462260723ULL;
            u32 xorshifted = (u32)(((state >> 18) ^ state) >> 27);
            u32 rot = (u32)(state >> 59);
            key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        }
        counter = 0;
        index = 64;  // empty
    }
    void refill() {
        for (int b = 0; b < 4; b++) chacha_block(key, counter + b, 0, 12, buf + 16 * b);
        counter += 4;
    }
    u32 next_u32() {
        if (index >= 64) { refill(); index = 0; }
        return buf[index++];
    }
    // rand_core BlockRng::next_u64 (three cases on the buffer position).
    u64 next_u64() {
        if (index < 63) {
            u64 v = ((u64)buf[index + 1] << 32) | buf[index];
            index += 2;
            return v;
        } else if (index >= 64) {
            refill();
            index = 2;
            return ((u64)buf[1] << 32) | buf[0];
        } else {
            u64 lo = buf[63];
            refill();
            index = 1;
            return ((u64)buf[0] << 32) | lo;
        }
    }
    // gen_range(0..n) for usize: UniformInt<usize>::sample_single_inclusive,
    // widening multiply + conservative zone rejection on u64 draws.
    u64 gen_range_usize(u64 n) {
        u64 range = n;  // (high-1) - low + 1
        if (range == 0) return next_u64();
        u64 zone = (range << __builtin_clzll(range)) - 1;
        for (;;) {
            u64 v = next_u64();
            unsigned __int128 m = (unsigned __int128)v * range;
            u64 hi = (u64)(m >> 64), lo = (u64)m;
            if (lo <= zone) return hi;
        }
    }
    u32 gen_range_u32(u32 range) {
        if (range == 0) return next_u32();
        u32 zone = (range << __builtin_clz(range)) - 1;
        for (;;) {
            u32 v = next_u32();
            u64 m = (u64)v * range;
            u32 hi = (u32)(m >> 32), lo = (u32)m;
            if (lo <= zone) return hi;
        }
    }
    // rand::seq::gen_index: u32 sampling when the bound fits in u32.
    u64 gen_index(u64 ubound) {
        if (ubound <= 0xFFFFFFFFull) return gen_range_u32((u32)ubound);
        return gen_range_usize(ubound);
    }
    // [0,1) float with 23 random mantissa bits from one u32 draw.
    float unit_f32() {
        u32 bits = (next_u32() >> 9) | 0x3F800000u;
        float f;
        memcpy(&f, &bits, 4);
        return f - 1.0f;
    }
    // gen_range(lo..hi) for f32: UniformFloat::sample_single.
    float gen_range_f32(float lo, float hi) {
        float scale = hi - lo;
        for (;;) {
            float res = unit_f32() * scale + lo;
            if (res < hi) return res;
        }
    }
};

// SliceRandom::shuffle: descending Fisher-Yates (kmeans.rs:258, :724).
template <class T>
static void rng_shuffle(StdRng& rng, std::vector<T>& v) {
    for (size_t i = v.size(); i-- > 1;) std::swap(v[i], v[rng.gen_index(i + 1)]);
}

// IteratorRandom::choose_multiple over 0..n: reservoir (kmeans.rs:595).
static std::vector<u64> rng_choose_multiple(StdRng& rng, u64 n, u64 amount) {
    std::vector<u64> res;
    u64 i = 0;
    for (; i < n && res.size() < amount; i++) res.push_back(i);
    if (res.size() == amount) {
        for (u64 j = 0; i < n; i++, j++) {
            u64 k = rng.gen_index(j + 1 + amount);
            if (k < amount) res[k] = i;
        }
    }
    return res;
}

// WeightedIndex<f32>::new + sample (kmeans.rs:204-206, :285-287).  Caller has
// already checked total != 0 (kmeans.rs:193-194).
static u64 weighted_index_sample(StdRng& rng, const std::vector<float>& w) {
    size_t n = w.size();
    std::vector<float> cum;
    cum.reserve(n ? n - 1 : 0);
    float total = w[0];
    for (size_t i = 1; i < n; i++) {
        cum.push_back(total);
        total += w[i];
    }
    // UniformFloat::new(0, total)
    const float max_rand = 1.0f - 1.1920929e-7f;  // 1 - 2^-23
    float scale = total - 0.0f;
    for (;;) {
        if (!(scale * max_rand + 0.0f >= total)) break;
        u32 b;
        memcpy(&b, &scale, 4);
        b -= 1;
        memcpy(&scale, &b, 4);
    }
    float x = rng.unit_f32() * scale + 0.0f;
    // binary_search_by(|w| if *w <= x {Less} else {Greater}).unwrap_err()
    size_t lo = 0, hi = cum.size();
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (cum[mid] <= x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ===========================================================================
// src/utils.rs
// ===========================================================================
// utils.rs:9-16
static u64 calculate_num_clusters(u64 n) {
    if (n < 10000) return (u64)std::sqrt((double)n);
    if (n < 100000) return 2 * (u64)std::ceil(std::sqrt((double)n));
    return 4 * (u64)std::ceil(std::sqrt((double)n));
}
// utils.rs:18-26
static u64 calculate_max_iterations(u64 n) {
    if (n < 10000) return 300;
    if (n < 100000) return 100;
    if (n < 1000000) return 50;
    return 20;
}
// utils.rs:28-30 -- strictly sequential left fold of (x-y)*(x-y), no FMA.
static float euclidean_distance_squared(const float* a, const float* b, size_t d) {
    float s = 0.0f;
    for (size_t i = 0; i < d; i++) {
        float diff = a[i] - b[i];
        s += diff * diff;
    }
    return s;
}

// ===========================================================================
// src/kmeans.rs
// ===========================================================================
// kmeans.rs:377-419 + wide 0.7.33 non-AVX reduce_add.
static float compute_distance_simd(const float* p, const float* c, size_t dim) {
    size_t j = 0;
    float a8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (j + 8 <= dim) {
        for (int l = 0; l < 8; l++) {
            float diff = p[j + l] - c[j + l];
            a8[l] += diff * diff;
        }
        j += 8;
    }
    float a4[4] = {0, 0, 0, 0};
    while (j + 4 <= dim) {
        for (int l = 0; l < 4; l++) {
            float diff = p[j + l] - c[j + l];
            a4[l] += diff * diff;
        }
        j += 4;
    }
    float tail = 0.0f;
    while (j < dim) {
        float diff = p[j] - c[j];
        tail += diff * diff;
        j++;
    }
    float lo = ((a8[0] + a8[1]) + a8[2]) + a8[3];
    float hi = ((a8[4] + a8[5]) + a8[6]) + a8[7];
    float r8 = lo + hi;
    float r4 = ((a4[0] + a4[1]) + a4[2]) + a4[3];
    return (r8 + r4) + tail;
}

// kmeans.rs:355-373 -- strict '<', first minimum wins.
static void find_nearest_centroid(const float* point, const float* cents, size_t k, size_t dim,
                                  size_t* best_c, float* best_d) {
    size_t bc = 0;
    float bd = std::numeric_limits<float>::infinity();
    for (size_t i = 0; i < k; i++) {
        float dist = compute_distance_simd(point, cents + i * dim, dim);
        if (dist < bd) { bd = dist; bc = i; }
    }
    *best_c = bc;
    *best_d = bd;
}

// kmeans.rs:422-443
static void update_min_distances(const float* data, size_t dim, const float* latest,
                                 std::vector<float>& min_d) {
    long m = (long)min_d.size();
#pragma omp parallel for schedule(static)
    for (long i = 0; i < m; i++) {
        float dist = compute_distance_simd(data + (size_t)i * dim, latest, dim);
        if (dist < min_d[i]) min_d[i] = dist;
    }
}

// kmeans.rs:154-310 (exact and sampled variants share everything but the sample map).
static void kmeans_plus_plus_init(const float* data, size_t n, size_t dim, size_t k, u64 seed,
                                  std::vector<float>& cents, std::vector<u64>* chosen_out) {
    StdRng rng(seed);
    const size_t sample_threshold = 50000;
    bool sampled = n > sample_threshold;
    size_t actual_k = std::min(k, n);
    cents.assign(k * dim, 0.0f);
    if (chosen_out) chosen_out->clear();

    u64 first = rng.gen_range_usize(n);
    memcpy(&cents[0], data + first * dim, dim * sizeof(float));
    if (chosen_out) chosen_out->push_back(first);

    std::vector<u64> sample_indices;
    size_t m = n;
    if (sampled) {
        sample_indices.resize(n);
        for (size_t i = 0; i < n; i++) sample_indices[i] = i;
        rng_shuffle(rng, sample_indices);
        sample_indices.resize(std::min(sample_threshold, n));
        m = sample_indices.size();
    }
    std::vector<float> min_d(m, std::numeric_limits<float>::infinity());
    std::vector<float> weights(m);
    for (size_t i = 1; i < actual_k; i++) {
        // NB (kmeans.rs:268, :431-436): in the sampled path distances are computed
        // for rows 0..m of `data`, not for the sampled rows.
        update_min_distances(data, dim, &cents[(i - 1) * dim], min_d);
        float total = 0.0f;
        for (size_t t = 0; t < m; t++) {
            weights[t] = min_d[t] * min_d[t];
            total += weights[t];
        }
        if (total == 0.0f) {
            u64 dup = rng.gen_range_usize(i);
            memcpy(&cents[i * dim], &cents[dup * dim], dim * sizeof(float));
            if (chosen_out) chosen_out->push_back((u64)-1 - dup);
        } else {
            u64 s = weighted_index_sample(rng, weights);
            u64 chosen = sampled ? sample_indices[s] : s;
            memcpy(&cents[i * dim], data + chosen * dim, dim * sizeof(float));
            if (chosen_out) chosen_out->push_back(chosen);
        }
    }
    for (size_t i = actual_k; i < k; i++) {
        u64 dup = rng.gen_range_usize(actual_k);
        memcpy(&cents[i * dim], &cents[dup * dim], dim * sizeof(float));
        if (chosen_out) chosen_out->push_back((u64)-1 - dup);
    }
}

// kmeans.rs:313-331
static void handle_empty_clusters(std::vector<float>& cents, const std::vector<u64>& counts,
                                  const float* data, size_t n, size_t dim, StdRng& rng) {
    size_t k = counts.size();
    for (size_t c = 0; c < k; c++) {
        if (counts[c] == 0) {
            u64 ri = rng.gen_range_usize(n);
            memcpy(&cents[c * dim], data + ri * dim, dim * sizeof(float));
        }
    }
}

// kmeans.rs:334-351.  The reference's rayon sum() has no defined order; the oracle
// fixes it to ascending centroid index (documented tolerance item).
static float compute_centroid_delta(const std::vector<float>& curr, const std::vector<float>& prev,
                                    size_t k, size_t dim) {
    float total = 0.0f;
    for (size_t c = 0; c < k; c++) {
        float local = 0.0f;
        for (size_t d = 0; d < dim; d++) {
            float diff = curr[c * dim + d] - prev[c * dim + d];
            local += diff * diff;
        }
        total += local;
    }
    return std::sqrt(total / (float)(k * dim));
}

// kmeans.rs:584-648
static void build_centroid_hierarchy(const float* cents, size_t k, size_t dim, size_t meta_k, u64 seed,
                                     std::vector<float>& meta, std::vector<u64>& c2m) {
    StdRng rng(seed);
    meta.assign(meta_k * dim, 0.0f);
    std::vector<u64> chosen = rng_choose_multiple(rng, k, meta_k);
    for (size_t i = 0; i < chosen.size(); i++)
        memcpy(&meta[i * dim], cents + chosen[i] * dim, dim * sizeof(float));
    c2m.assign(k, 0);
    for (int iter = 0; iter < 5; iter++) {
        for (size_t c = 0; c < k; c++) {
            size_t best = 0;
            float bd = std::numeric_limits<float>::infinity();
            for (size_t m = 0; m < meta_k; m++) {
                float dist = compute_distance_simd(cents + c * dim, &meta[m * dim], dim);
                if (dist < bd) { bd = dist; best = m; }
            }
            c2m[c] = best;
        }
        std::vector<float> sum(dim);
        for (size_t m = 0; m < meta_k; m++) {
            size_t count = 0;
            std::fill(sum.begin(), sum.end(), 0.0f);
            for (size_t c = 0; c < k; c++) {
                if (c2m[c] == m) {
                    count++;
                    for (size_t d = 0; d < dim; d++) sum[d] += cents[c * dim + d];
                }
            }
            if (count > 0)
                for (size_t d = 0; d < dim; d++) meta[m * dim + d] = sum[d] / (float)count;
        }
    }
}

// kmeans.rs:474-581 (+ :651-672 for the stable top-3 meta sort)
static void assign_points_hierarchical(const float* data, size_t n, size_t dim, const float* cents,
                                       size_t k, u64 seed, u64* labels) {
    size_t meta_k = std::min(std::max((size_t)std::sqrt((float)k), (size_t)2), k / 2);
    u64 hseed = seed * 17ULL + 42ULL;
    std::vector<float> meta;
    std::vector<u64> c2m;
    build_centroid_hierarchy(cents, k, dim, meta_k, hseed, meta, c2m);
    std::vector<std::vector<u64>> m2c(meta_k);
    for (size_t c = 0; c < k; c++) m2c[c2m[c]].push_back(c);
    size_t top = std::min((size_t)3, meta_k);
#pragma omp parallel
    {
        std::vector<std::pair<float, u64>> dist(meta_k);
#pragma omp for schedule(static)
        for (long i = 0; i < (long)n; i++) {
            const float* p = data + (size_t)i * dim;
            for (size_t m = 0; m < meta_k; m++)
                dist[m] = std::make_pair(compute_distance_simd(p, &meta[m * dim], dim), (u64)m);
            std::stable_sort(dist.begin(), dist.end(),
                             [](const std::pair<float, u64>& a, const std::pair<float, u64>& b) {
                                 return a.first < b.first;
                             });
            u64 best = 0;
            bool any = false;
            float bd = std::numeric_limits<float>::infinity();
            for (size_t t = 0; t < top; t++) {
                for (u64 c : m2c[dist[t].second]) {
                    float dd = compute_distance_simd(p, cents + c * dim, dim);
                    if (!any) { best = c; any = true; }  // candidate_indices[0] default
                    if (dd < bd) { bd = dd; best = c; }
                }
            }
            labels[i] = best;
        }
    }
}

// kmeans.rs:462-470
static void assign_points_brute_force(const float* data, size_t n, size_t dim, const float* cents,
                                      size_t k, u64* labels) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        size_t bc;
        float bd;
        find_nearest_centroid(data + (size_t)i * dim, cents, k, dim, &bc, &bd);
        labels[i] = bc;
    }
}

// kmeans.rs:445-459
static void assign_points(const float* data, size_t n, size_t dim, const float* cents, size_t k,
                          u64 seed, u64* labels) {
    if (k > 100) assign_points_hierarchical(data, n, dim, cents, k, seed, labels);
    else assign_points_brute_force(data, n, dim, cents, k, labels);
}

// kmeans.rs:674-719
static void update_centroids_full(const float* data, size_t n, size_t dim, const u64* labels, size_t k,
                                  std::vector<float>& out, std::vector<u64>& counts) {
    out.assign(k * dim, 0.0f);
    counts.assign(k, 0);
    for (size_t i = 0; i < n; i++) {
        size_t c = labels[i];
        counts[c]++;
        for (size_t d = 0; d < dim; d++) out[c * dim + d] += data[i * dim + d];
    }
    for (size_t c = 0; c < k; c++)
        if (counts[c] > 0)
            for (size_t d = 0; d < dim; d++) out[c * dim + d] /= (float)counts[c];
}

// kmeans.rs:729-787
static void update_centroids_mini_batch(const float* data, size_t dim, std::vector<float>& cents,
                                        const std::vector<u64>& batch, const std::vector<u64>& blabels,
                                        std::vector<u64>& counts) {
    size_t k = counts.size();
    std::vector<std::vector<u64>> pts(k);
    for (size_t t = 0; t < batch.size(); t++) pts[blabels[t]].push_back(batch[t]);
    std::vector<float> sum(dim);
    for (size_t c = 0; c < k; c++) {
        if (pts[c].empty()) continue;
        u64 new_count = counts[c] + 1;  // +1 per batch, not per point (kmeans.rs:757)
        float eta = 1.0f / (float)new_count;
        std::fill(sum.begin(), sum.end(), 0.0f);
        for (u64 idx : pts[c])
            for (size_t d = 0; d < dim; d++) sum[d] += data[idx * dim + d];
        for (size_t d = 0; d < dim; d++) {
            float mean = sum[d] / (float)pts[c].size();
            cents[c * dim + d] = (1.0f - eta) * cents[c * dim + d] + eta * mean;
        }
        counts[c] = new_count;
    }
}

// kmeans.rs:64-150
static int run_kmeans_mini_batch(const float* data, size_t n, size_t dim, size_t k, size_t max_iters,
                                 float tol, u64 seed, std::vector<float>& cents, u64* labels,
                                 u64* iters_run) {
    if (n == 0 || dim == 0) return 1;  // ErrorKind::InvalidInput (kmeans.rs:72-77)
    if (tol < 0) tol = 1e-4f;
    StdRng rng(seed);
    size_t batch_size = std::min((size_t)256, std::max((size_t)10, (size_t)std::sqrt((float)n)));
    kmeans_plus_plus_init(data, n, dim, k, seed, cents, nullptr);
    std::vector<u64> counts(k, 0);
    std::vector<float> prev = cents;
    std::vector<u64> idx(n);
    u64 it = 0;
    for (; it < max_iters;) {
        // sample_batch (kmeans.rs:722-726): shuffle all n, take batch_size
        for (size_t i = 0; i < n; i++) idx[i] = i;
        rng_shuffle(rng, idx);
        size_t b = std::min(batch_size, n);
        std::vector<u64> batch(idx.begin(), idx.begin() + b), bl(b);
#pragma omp parallel for schedule(static)
        for (long t = 0; t < (long)b; t++) {
            size_t bc;
            float bd;
            find_nearest_centroid(data + batch[t] * dim, cents.data(), k, dim, &bc, &bd);
            bl[t] = bc;
        }
        update_centroids_mini_batch(data, dim, cents, batch, bl, counts);
        handle_empty_clusters(cents, counts, data, n, dim, rng);
        float delta = compute_centroid_delta(cents, prev, k, dim);
        prev = cents;
        it++;
        if (delta < tol) break;
    }
    if (iters_run) *iters_run = it;
    assign_points(data, n, dim, cents.data(), k, seed, labels);
    return 0;
}

// kmeans.rs:15-60
static int run_kmeans_parallel(const float* data, size_t n, size_t dim, size_t k, size_t max_iters,
                               float tol, u64 seed, std::vector<float>& cents, u64* labels,
                               u64* iters_run) {
    if (n == 0 || dim == 0) return 1;
    if (tol < 0) tol = 1e-4f;
    StdRng rng(seed);
    kmeans_plus_plus_init(data, n, dim, k, seed, cents, nullptr);
    for (size_t i = 0; i < n; i++) labels[i] = 0;
    u64 it = 0;
    for (; it < max_iters;) {
        assign_points(data, n, dim, cents.data(), k, seed, labels);
        std::vector<float> nc;
        std::vector<u64> counts;
        update_centroids_full(data, n, dim, labels, k, nc, counts);
        handle_empty_clusters(nc, counts, data, n, dim, rng);
        float delta = compute_centroid_delta(nc, cents, k, dim);
        cents = nc;
        it++;
        if (delta < tol) break;
    }
    if (iters_run) *iters_run = it;
    return 0;
}

// ===========================================================================
// src/ivf_index.rs  (in-memory lists; the shard files are a separate codec below)
// ===========================================================================
struct OracleIvf {
    u32 dim = 0;
    size_t n = 0;
    std::vector<float> data;                 // n x dim (vector_store.rs:48-58)
    std::vector<u64> ext_id, ts;             // per internal id (= row index)
    std::vector<float> centroids;            // nlist x dim, non-empty lists only (ivf_index.rs:142-146)
    std::vector<u64> c2shard;                // ivf_index.rs:149-164
    std::vector<std::vector<u64>> lists;     // internal ids, ascending (ivf_index.rs:94-101)
    std::vector<u64> labels_all;             // label of every vector in the UNfiltered numbering
    std::vector<float> centroids_all;        // k x dim before the empty-list filter
    std::vector<u64> old_to_new;             // k entries, (u64)-1 for dropped lists
    u64 num_shards = 0, k_trained = 0, iters_run = 0;
    std::vector<unsigned char> list_mask;  // optional: lists scanned by this "rank" (multi-GPU tests)
};

// ivf_index.rs:58-177
static OracleIvf* ivf_fit(const float* data, const u64* ext, const u64* ts, size_t n, size_t dim,
                          u64 seed, u64 nlist_override, u64 iters_override) {
    OracleIvf* ix = new OracleIvf();
    ix->dim = (u32)dim;
    ix->n = n;
    ix->data.assign(data, data + n * dim);
    ix->ext_id.resize(n);
    ix->ts.resize(n);
    for (size_t i = 0; i < n; i++) {
        ix->ext_id[i] = ext ? ext[i] : i;
        ix->ts[i] = ts ? ts[i] : 0;
    }
    u64 k = nlist_override ? nlist_override : calculate_num_clusters(n);
    u64 max_iters = iters_override ? iters_override : calculate_max_iterations(n);
    ix->k_trained = k;
    ix->labels_all.resize(n);
    run_kmeans_mini_batch(data, n, dim, k, max_iters, -1.0f, seed, ix->centroids_all,
                          ix->labels_all.data(), &ix->iters_run);
    std::vector<std::vector<u64>> lists(k);
    for (size_t i = 0; i < n; i++) lists[ix->labels_all[i]].push_back(i);
    u64 num_shards = (u64)std::ceil(std::sqrt((float)k));
    u64 super_seed = seed * 31ULL + 7ULL;
    std::vector<float> super_c;
    std::vector<u64> super_labels(k);
    run_kmeans_mini_batch(ix->centroids_all.data(), k, dim, num_shards, 100, -1.0f, super_seed, super_c,
                          super_labels.data(), nullptr);
    ix->num_shards = num_shards;
    ix->old_to_new.assign(k, (u64)-1);
    for (size_t c = 0; c < k; c++) {
        if (lists[c].empty()) continue;
        ix->old_to_new[c] = ix->lists.size();
        ix->centroids.insert(ix->centroids.end(), ix->centroids_all.begin() + c * dim,
                             ix->centroids_all.begin() + (c + 1) * dim);
        ix->c2shard.push_back(super_labels[c]);
        ix->lists.push_back(std::move(lists[c]));
    }
    return ix;
}

// ivf_index.rs:190-267 for one query.  Candidate order = probe-rank order, then
// list order (the reference groups by shard in HashSet order, which is
// process-random; ties across lists are "unspecified" there).  Returns count.
static long ivf_search_one(const OracleIvf* ix, const float* q, size_t k, size_t nprobe, u64* out_id,
                           float* out_d, u64* out_internal,
                           std::vector<std::pair<float, u64>>& cd, std::vector<std::pair<float, u64>>& cand) {
    if (k == 0 || nprobe == 0) return -1;  // InvalidInput (ivf_index.rs:197-202)
    size_t nlist = ix->lists.size(), dim = ix->dim;
    cd.resize(nlist);
    for (size_t c = 0; c < nlist; c++) {
        float d = euclidean_distance_squared(q, &ix->centroids[c * dim], dim);
        if (d != d) return -2;  // partial_cmp().unwrap() panics on NaN
        cd[c] = std::make_pair(d, (u64)c);
    }
    auto by_dist = [](const std::pair<float, u64>& a, const std::pair<float, u64>& b) { return a.first < b.first; };
    std::stable_sort(cd.begin(), cd.end(), by_dist);
    size_t np = std::min(nprobe, nlist);
    cand.clear();
    for (size_t r = 0; r < np; r++) {
        if (!ix->list_mask.empty() && !ix->list_mask[cd[r].second]) continue;
        for (u64 id : ix->lists[cd[r].second]) {
            float d = euclidean_distance_squared(q, &ix->data[id * dim], dim);
            if (d != d) return -2;
            cand.push_back(std::make_pair(d, id));
        }
    }
    std::stable_sort(cand.begin(), cand.end(), by_dist);
    size_t m = std::min(k, cand.size());
    for (size_t i = 0; i < m; i++) {
        out_d[i] = cand[i].first;
        out_id[i] = ix->ext_id[cand[i].second];
        if (out_internal) out_internal[i] = cand[i].second;
    }
    return (long)m;
}

// ===========================================================================
// src/shards.rs file codec (header 40 B, index entries 32 B, AoS blocks)
// ===========================================================================
#pragma pack(push, 1)
struct ShardHeader { u64 shard_id, version; u32 dimensions, num_centroids; u64 index_offset, data_offset; };
struct CentroidIndex { u64 centroid_id; u32 num_vectors, pad; u64 data_offset, data_size; };
struct VectorMeta { u64 id, external_id, timestamp; };
#pragma pack(pop)
static_assert(sizeof(ShardHeader) == 40, "shards.rs:22-31 repr(C) is 40 bytes");
static_assert(sizeof(CentroidIndex) == 32, "shards.rs:34-42");
static_assert(sizeof(VectorMeta) == 24, "shards.rs:45-51");

// ===========================================================================
// C ABI for ctypes (tests / bench only)
// ===========================================================================
extern "C" {

void vo_chacha_block(const u32* key, u64 counter, u64 stream, int rounds, u32* out) {
    chacha_block(key, counter, stream, rounds, out);
}
void* vo_rng_new(u64 seed) { return new StdRng(seed); }
void vo_rng_free(void* r) { delete (StdRng*)r; }
void vo_rng_key(void* r, u32* out) { memcpy(out, ((StdRng*)r)->key, 32); }
u32 vo_rng_next_u32(void* r) { return ((StdRng*)r)->next_u32(); }
u64 vo_rng_next_u64(void* r) { return ((StdRng*)r)->next_u64(); }
u64 vo_rng_gen_range(void* r, u64 n) { return ((StdRng*)r)->gen_range_usize(n); }
u64 vo_rng_gen_index(void* r, u64 n) { return ((StdRng*)r)->gen_index(n); }
float vo_rng_gen_range_f32(void* r, float lo, float hi) { return ((StdRng*)r)->gen_range_f32(lo, hi); }
void vo_rng_shuffle(void* r, u64* arr, u64 n) {
    std::vector<u64> v(arr, arr + n);
    rng_shuffle(*(StdRng*)r, v);
    memcpy(arr, v.data(), n * sizeof(u64));
}
u64 vo_rng_choose_multiple(void* r, u64 n, u64 amount, u64* out) {
    std::vector<u64> v = rng_choose_multiple(*(StdRng*)r, n, amount);
    memcpy(out, v.data(), v.size() * sizeof(u64));
    return v.size();
}
u64 vo_rng_weighted_index(void* r, const float* w, u64 n) {
    std::vector<float> v(w, w + n);
    return weighted_index_sample(*(StdRng*)r, v);
}
// tests/test_utils/mod.rs:245-252 create_deterministic_vectors
void vo_create_deterministic_vectors(u64 n, u64 dim, u64 seed, float* out) {
    StdRng rng(seed);
    for (u64 i = 0; i < n * dim; i++) out[i] = rng.gen_range_f32(-10.0f, 10.0f);
}

u64 vo_calculate_num_clusters(u64 n) { return calculate_num_clusters(n); }
u64 vo_calculate_max_iterations(u64 n) { return calculate_max_iterations(n); }
float vo_euclidean_distance_squared(const float* a, const float* b, u64 d) {
    return euclidean_distance_squared(a, b, d);
}
float vo_compute_distance_simd(const float* a, const float* b, u64 d) { return compute_distance_simd(a, b, d); }

void vo_kmeans_pp_init(const float* data, u64 n, u64 dim, u64 k, u64 seed, float* out_c, u64* out_chosen) {
    std::vector<float> c;
    std::vector<u64> ch;
    kmeans_plus_plus_init(data, n, dim, k, seed, c, &ch);
    memcpy(out_c, c.data(), c.size() * sizeof(float));
    if (out_chosen) memcpy(out_chosen, ch.data(), ch.size() * sizeof(u64));
}
int vo_kmeans_mini_batch(const float* data, u64 n, u64 dim, u64 k, u64 max_iters, float tol, u64 seed,
                         float* out_c, u64* out_labels, u64* iters_run) {
    std::vector<float> c;
    int rc = run_kmeans_mini_batch(data, n, dim, k, max_iters, tol, seed, c, out_labels, iters_run);
    if (rc == 0) memcpy(out_c, c.data(), c.size() * sizeof(float));
    return rc;
}
int vo_kmeans_parallel(const float* data, u64 n, u64 dim, u64 k, u64 max_iters, float tol, u64 seed,
                       float* out_c, u64* out_labels, u64* iters_run) {
    std::vector<float> c;
    int rc = run_kmeans_parallel(data, n, dim, k, max_iters, tol, seed, c, out_labels, iters_run);
    if (rc == 0) memcpy(out_c, c.data(), c.size() * sizeof(float));
    return rc;
}
void vo_assign_points(const float* data, u64 n, u64 dim, const float* cents, u64 k, u64 seed, u64* labels) {
    assign_points(data, n, dim, cents, k, seed, labels);
}
void vo_assign_brute_force(const float* data, u64 n, u64 dim, const float* cents, u64 k, u64* labels) {
    assign_points_brute_force(data, n, dim, cents, k, labels);
}
u64 vo_build_hierarchy(const float* cents, u64 k, u64 dim, u64 seed, float* out_meta, u64* out_c2m) {
    size_t meta_k = std::min(std::max((size_t)std::sqrt((float)k), (size_t)2), (size_t)(k / 2));
    std::vector<float> meta;
    std::vector<u64> c2m;
    build_centroid_hierarchy(cents, k, dim, meta_k, seed * 17ULL + 42ULL, meta, c2m);
    if (out_meta) memcpy(out_meta, meta.data(), meta.size() * sizeof(float));
    if (out_c2m) memcpy(out_c2m, c2m.data(), c2m.size() * sizeof(u64));
    return meta_k;
}
void vo_update_centroids_full(const float* data, u64 n, u64 dim, const u64* labels, u64 k, float* out_c,
                              u64* out_counts) {
    std::vector<float> c;
    std::vector<u64> cnt;
    update_centroids_full(data, n, dim, labels, k, c, cnt);
    memcpy(out_c, c.data(), c.size() * sizeof(float));
    memcpy(out_counts, cnt.data(), cnt.size() * sizeof(u64));
}
float vo_centroid_delta(const float* a, const float* b, u64 k, u64 dim) {
    std::vector<float> x(a, a + k * dim), y(b, b + k * dim);
    return compute_centroid_delta(x, y, k, dim);
}

void* vo_ivf_fit(const float* data, const u64* ext, const u64* ts, u64 n, u64 dim, u64 seed, u64 nlist_override,
                 u64 iters_override) {
    if (n == 0 || dim == 0) return nullptr;
    return ivf_fit(data, ext, ts, n, dim, seed, nlist_override, iters_override);
}
// Build from externally supplied centroids/labels (used to cross-check the scan
// alone, decoupled from training).
void* vo_ivf_from_labels(const float* data, const u64* ext, u64 n, u64 dim, const float* cents, u64 k,
                         const u64* labels) {
    OracleIvf* ix = new OracleIvf();
    ix->dim = (u32)dim;
    ix->n = n;
    ix->data.assign(data, data + n * dim);
    ix->ext_id.resize(n);
    ix->ts.assign(n, 0);
    for (size_t i = 0; i < n; i++) ix->ext_id[i] = ext ? ext[i] : i;
    std::vector<std::vector<u64>> lists(k);
    for (size_t i = 0; i < n; i++) lists[labels[i]].push_back(i);
    ix->k_trained = k;
    ix->old_to_new.assign(k, (u64)-1);
    for (size_t c = 0; c < k; c++) {
        if (lists[c].empty()) continue;
        ix->old_to_new[c] = ix->lists.size();
        ix->centroids.insert(ix->centroids.end(), cents + c * dim, cents + (c + 1) * dim);
        ix->c2shard.push_back(0);
        ix->lists.push_back(std::move(lists[c]));
    }
    return ix;
}
void vo_ivf_free(void* h) { delete (OracleIvf*)h; }
// Restrict the scan to a subset of lists (probe selection still sees every centroid):
// what one rank of the sharded multi-GPU search does.  mask == NULL clears it.
void vo_ivf_set_list_mask(void* h, const unsigned char* mask) {
    OracleIvf* ix = (OracleIvf*)h;
    if (mask) ix->list_mask.assign(mask, mask + ix->lists.size()); else ix->list_mask.clear();
}
u64 vo_ivf_nlist(void* h) { return ((OracleIvf*)h)->lists.size(); }
u64 vo_ivf_k_trained(void* h) { return ((OracleIvf*)h)->k_trained; }
u64 vo_ivf_num_shards(void* h) { return ((OracleIvf*)h)->num_shards; }
u64 vo_ivf_iters_run(void* h) { return ((OracleIvf*)h)->iters_run; }
void vo_ivf_centroids(void* h, float* out) {
    OracleIvf* ix = (OracleIvf*)h;
    memcpy(out, ix->centroids.data(), ix->centroids.size() * sizeof(float));
}
void vo_ivf_centroids_all(void* h, float* out) {
    OracleIvf* ix = (OracleIvf*)h;
    memcpy(out, ix->centroids_all.data(), ix->centroids_all.size() * sizeof(float));
}
void vo_ivf_labels_all(void* h, u64* out) {
    OracleIvf* ix = (OracleIvf*)h;
    memcpy(out, ix->labels_all.data(), ix->labels_all.size() * sizeof(u64));
}
void vo_ivf_c2shard(void* h, u64* out) {
    OracleIvf* ix = (OracleIvf*)h;
    memcpy(out, ix->c2shard.data(), ix->c2shard.size() * sizeof(u64));
}
void vo_ivf_list_sizes(void* h, u64* out) {
    OracleIvf* ix = (OracleIvf*)h;
    for (size_t l = 0; l < ix->lists.size(); l++) out[l] = ix->lists[l].size();
}
void vo_ivf_list_members(void* h, u64 l, u64* out) {
    OracleIvf* ix = (OracleIvf*)h;
    memcpy(out, ix->lists[l].data(), ix->lists[l].size() * sizeof(u64));
}
// Single query, reference semantics.  Returns number of results (<= k), -1 for
// InvalidInput, -2 where the reference would panic on NaN.
long vo_ivf_search(void* h, const float* q, u64 k, u64 nprobe, u64* out_id, float* out_d) {
    std::vector<std::pair<float, u64>> a, b;
    return ivf_search_one((OracleIvf*)h, q, k, nprobe, out_id, out_d, nullptr, a, b);
}
// Probe list of one query (coarse stage only): ivf_index.rs:205-220.
long vo_ivf_probes(void* h, const float* q, u64 nprobe, u64* out_list, float* out_d) {
    OracleIvf* ix = (OracleIvf*)h;
    size_t nlist = ix->lists.size(), dim = ix->dim;
    std::vector<std::pair<float, u64>> cd(nlist);
    for (size_t c = 0; c < nlist; c++)
        cd[c] = std::make_pair(euclidean_distance_squared(q, &ix->centroids[c * dim], dim), (u64)c);
    std::stable_sort(cd.begin(), cd.end(),
                     [](const std::pair<float, u64>& a, const std::pair<float, u64>& b) { return a.first < b.first; });
    size_t np = std::min((size_t)nprobe, nlist);
    for (size_t r = 0; r < np; r++) { out_list[r] = cd[r].second; out_d[r] = cd[r].first; }
    return (long)np;
}
// The PyO3 batched boundary (bindings/python/src/lib.rs:123-203): D init +inf, I init -1.
// nthreads = 1 is the reference's behaviour (queries strictly sequential,
// lib.rs:74-97); nthreads > 1 is the generous all-core baseline.
int vo_ivf_search_batch(void* h, const float* xq, u64 nq, u64 k, u64 nprobe, float* D, int64_t* I, int nthreads) {
    OracleIvf* ix = (OracleIvf*)h;
    if (k == 0 || nprobe == 0) return -1;
    size_t dim = ix->dim;
    int bad = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        std::vector<std::pair<float, u64>> a, b;
        std::vector<u64> ids(k);
        std::vector<float> ds(k);
#pragma omp for schedule(dynamic, 4)
        for (long qi = 0; qi < (long)nq; qi++) {
            for (size_t t = 0; t < k; t++) {
                D[qi * k + t] = std::numeric_limits<float>::infinity();
                I[qi * k + t] = -1;
            }
            long m = ivf_search_one(ix, xq + (size_t)qi * dim, k, nprobe, ids.data(), ds.data(), nullptr, a, b);
            if (m < 0) { bad = 1; continue; }
            for (long t = 0; t < m; t++) {
                D[qi * k + t] = ds[t];
                I[qi * k + t] = (int64_t)ids[t];
            }
        }
    }
    return bad ? -2 : 0;
}

// Exact brute-force top-k (tests/test_utils/mod.rs:225-235 find_true_nearest_neighbors),
// all cores; used as ground truth for recall.
void vo_brute_force_topk(const float* data, u64 n, u64 dim, const float* xq, u64 nq, u64 k, int64_t* I, float* D) {
#pragma omp parallel
    {
        std::vector<std::pair<float, u64>> d(n);
#pragma omp for schedule(dynamic, 1)
        for (long qi = 0; qi < (long)nq; qi++) {
            for (size_t i = 0; i < n; i++)
                d[i] = std::make_pair(euclidean_distance_squared(xq + (size_t)qi * dim, data + i * dim, dim), (u64)i);
            size_t m = std::min((size_t)k, (size_t)n);
            std::partial_sort(d.begin(), d.begin() + m, d.end());
            for (size_t t = 0; t < k; t++) {
                I[qi * k + t] = t < m ? (int64_t)d[t].second : -1;
                if (D) D[qi * k + t] = t < m ? d[t].first : std::numeric_limits<float>::infinity();
            }
        }
    }
}

// ---- shard codec (shards.rs:68-177 writer, :188-349 reader) -----------------
// Writes one shard file from flat arrays: for list i (0..nlists): centroid id,
// centroid vector, len, then per vector (id, external_id, timestamp, data).
int vo_shard_write(const char* path, u64 shard_id, u32 dim, u32 nlists, const u64* centroid_ids,
                   const float* centroid_vecs, const u32* lens, const u64* meta /*3 per vector*/,
                   const float* vecs) {
    FILE* f = fopen(path, "wb");
    if (!f) return 1;
    size_t vsz = (size_t)dim * 4, cpad = (8 - vsz % 8) % 8, vpad = cpad;
    ShardHeader hd = {shard_id, 1, dim, nlists, 40, 40 + 32ull * nlists};
    fwrite(&hd, sizeof(hd), 1, f);
    u64 off = hd.data_offset;
    for (u32 i = 0; i < nlists; i++) {
        u64 size = vsz + cpad + (u64)lens[i] * (24 + vsz + vpad);
        CentroidIndex e = {centroid_ids[i], lens[i], 0, off, size};
        fwrite(&e, sizeof(e), 1, f);
        off += size;
    }
    const char zeros[8] = {0};
    size_t v = 0;
    for (u32 i = 0; i < nlists; i++) {
        fwrite(centroid_vecs + (size_t)i * dim, 4, dim, f);
        fwrite(zeros, 1, cpad, f);
        for (u32 j = 0; j < lens[i]; j++, v++) {
            fwrite(meta + 3 * v, 8, 3, f);
            fwrite(vecs + v * dim, 4, dim, f);
            fwrite(zeros, 1, vpad, f);
        }
    }
    fclose(f);
    return 0;
}
// Reader: pass 1 (out arrays NULL) returns counts; pass 2 fills.  Return codes:
// 0 ok, 1 open failed (ErrorKind::Other), 2 invalid data / shard id mismatch.
int vo_shard_read(const char* path, u64 expect_shard_id, u32* dim, u32* nlists, u64* total_vectors, u64* centroid_ids,
                  float* centroid_vecs, u32* lens, u64* meta, float* vecs) {
    FILE* f = fopen(path, "rb");
    if (!f) return 1;
    ShardHeader hd;
    if (fread(&hd, sizeof(hd), 1, f) != 1) { fclose(f); return 2; }
    if (hd.shard_id != expect_shard_id) { fclose(f); return 2; }
    std::vector<CentroidIndex> idx(hd.num_centroids);
    fseek(f, (long)hd.index_offset, SEEK_SET);
    if (hd.num_centroids && fread(idx.data(), 32, hd.num_centroids, f) != hd.num_centroids) { fclose(f); return 2; }
    *dim = hd.dimensions;
    *nlists = hd.num_centroids;
    u64 tot = 0;
    for (auto& e : idx) tot += e.num_vectors;
    *total_vectors = tot;
    if (!centroid_ids) { fclose(f); return 0; }
    size_t d = hd.dimensions, vsz = d * 4, cpad = (8 - vsz % 8) % 8;
    size_t v = 0;
    for (u32 i = 0; i < hd.num_centroids; i++) {
        centroid_ids[i] = idx[i].centroid_id;
        lens[i] = idx[i].num_vectors;
        std::vector<char> blk(idx[i].data_size);
        fseek(f, (long)idx[i].data_offset, SEEK_SET);
        if (idx[i].data_size && fread(blk.data(), 1, blk.size(), f) != blk.size()) { fclose(f); return 2; }
        if (blk.size() < vsz) { fclose(f); return 2; }
        memcpy(centroid_vecs + (size_t)i * d, blk.data(), vsz);
        size_t off = vsz + cpad;
        for (u32 j = 0; j < idx[i].num_vectors; j++, v++) {
            if (off + 24 + vsz > blk.size()) { fclose(f); return 2; }
            memcpy(meta + 3 * v, blk.data() + off, 24);
            memcpy(vecs + v * d, blk.data() + off + 24, vsz);
            off += 24 + vsz + cpad;
        }
    }
    fclose(f);
    return 0;
}

int vo_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"

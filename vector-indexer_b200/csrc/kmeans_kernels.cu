// kmeans_kernels.cu -- k-means building blocks on sm_100a (src/kmeans.rs).
//
// Arithmetic contract: distances reproduce compute_distance_simd (kmeans.rs:377-419)
// bit for bit: eight strided lane sums over the 8-wide chunks (lane l accumulates dims
// 8c+l in chunk order, separate sub/mul/add roundings), at most one 4-wide chunk, a
// sequential tail, combined as
//   ((l0+l1)+l2)+l3  +  ((l4+l5)+l6)+l7 ,  + ((a0+a1)+a2)+a3 ,  + tail
// (wide 0.7.33 non-AVX reduce_add).  Argmin is strict '<', first minimum wins
// (kmeans.rs:355-373), realised as a 64-bit atomicMin on (distance bits << 32 | position):
// distances are non-negative so their bit patterns order like the values.
#include "kmeans.h"

namespace vidx {

constexpr int kPT = 32;        // points per tile
constexpr int kCT = 64;        // centroids per tile
constexpr int kDC = 32;        // dims per shared-memory chunk (4 sub-chunks of 8)
constexpr int kRowStride = kDC + 4;
constexpr int kPairsThreads = 256;
constexpr uint64_t kInitKey = (uint64_t)0x7f800000u << 32;  // (+inf, position 0): kmeans.rs:360-361

__device__ __forceinline__ float reduce8(const float (&a)[8]) {
    float lo = __fadd_rn(__fadd_rn(__fadd_rn(a[0], a[1]), a[2]), a[3]);
    float hi = __fadd_rn(__fadd_rn(__fadd_rn(a[4], a[5]), a[6]), a[7]);
    return __fadd_rn(lo, hi);
}

// Load rows [row0, row0+nrows) x dims [d0, d0+nd) of a gathered row set into smem
// (row stride kRowStride), zero filling absent rows / dims.
template <class RowFn>
__device__ __forceinline__ void load_rows(float* s, int nrows_tile, int nrows_valid, RowFn row_of, const float* base, int D,
                                          int d0, int nd, bool aligned) {
    int tid = threadIdx.x;
    if (aligned) {
        int nf4 = kDC / 4;
        for (int idx = tid; idx < nrows_tile * nf4; idx += kPairsThreads) {
            int r = idx / nf4, c = idx - r * nf4;
            float4 v = make_float4(0, 0, 0, 0);
            if (r < nrows_valid && 4 * c < nd) {
                const float* p = base + (size_t)row_of(r) * D + d0 + 4 * c;
                if (4 * c + 4 <= nd) v = __ldg(reinterpret_cast<const float4*>(p));
                else {
                    v.x = p[0];
                    if (4 * c + 1 < nd) v.y = p[1];
                    if (4 * c + 2 < nd) v.z = p[2];
                }
            }
            *reinterpret_cast<float4*>(&s[r * kRowStride + 4 * c]) = v;
        }
    } else {
        for (int idx = tid; idx < nrows_tile * kDC; idx += kPairsThreads) {
            int r = idx / kDC, c = idx - r * kDC;
            float v = 0.0f;
            if (r < nrows_valid && c < nd) v = base[(size_t)row_of(r) * D + d0 + c];
            s[r * kRowStride + c] = v;
        }
    }
}

// One block = one tile of <= 32 points of one item, looped over all centroid tiles of the
// item.  MODE 0: write distances out[(pt_off + i) * ld + j].  MODE 1: atomicMin keys.
template <int MODE>
__global__ void __launch_bounds__(kPairsThreads, 2)
pairs_kernel(const float* __restrict__ data, int D, const float* __restrict__ cents, const PairItem* __restrict__ items,
             const uint32_t* __restrict__ item_tile_off, int nitems, const uint2* __restrict__ pt_entries,
             const uint32_t* __restrict__ cent_idx, float* __restrict__ out, uint64_t ld,
             unsigned long long* __restrict__ best) {
    __shared__ __align__(16) float sp[kPT * kRowStride];
    __shared__ __align__(16) float sc[kCT * kRowStride];
    int tid = threadIdx.x;
    // locate the item of this block (item_tile_off has nitems+1 entries)
    int lo = 0, hi = nitems;
    uint32_t b = blockIdx.x;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (item_tile_off[mid] <= b) lo = mid; else hi = mid;
    }
    PairItem it = items[lo];
    uint32_t ptile = b - item_tile_off[lo];
    uint32_t p0 = ptile * kPT;
    int npv = (int)min((uint32_t)kPT, it.npts - p0);
    bool aligned = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(data) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(cents) & 15) == 0);
    int Dmain = D & ~7;
    int rem = D - Dmain;
    int tp = tid >> 4, tc = tid & 15;

    auto pt_row = [&](int r) -> uint32_t {
        uint32_t slot = it.pt_off + p0 + r;
        return pt_entries ? pt_entries[slot].x : slot;
    };
    for (uint32_t c0 = 0; c0 < it.ncents; c0 += kCT) {
        int ncv = (int)min((uint32_t)kCT, it.ncents - c0);
        auto c_row = [&](int r) -> uint32_t {
            uint32_t j = it.c_off + c0 + r;
            return cent_idx ? cent_idx[j] : j;
        };
        float acc[2][4][8];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int l = 0; l < 8; l++) acc[a][j][l] = 0.0f;

        for (int d0 = 0; d0 < Dmain; d0 += kDC) {
            int nd = min(kDC, Dmain - d0);
            __syncthreads();
            load_rows(sp, kPT, npv, pt_row, data, D, d0, nd, aligned);
            load_rows(sc, kCT, ncv, c_row, cents, D, d0, nd, aligned);
            __syncthreads();
            int nsub = nd >> 3;
            for (int sidx = 0; sidx < nsub; sidx++) {
                float4 pa[2][2], ca[4][2];
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    pa[a][0] = *reinterpret_cast<const float4*>(&sp[(2 * tp + a) * kRowStride + sidx * 8]);
                    pa[a][1] = *reinterpret_cast<const float4*>(&sp[(2 * tp + a) * kRowStride + sidx * 8 + 4]);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    ca[j][0] = *reinterpret_cast<const float4*>(&sc[(tc + 16 * j) * kRowStride + sidx * 8]);
                    ca[j][1] = *reinterpret_cast<const float4*>(&sc[(tc + 16 * j) * kRowStride + sidx * 8 + 4]);
                }
#pragma unroll
                for (int a = 0; a < 2; a++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        acc[a][j][0] = sqdiff_acc(acc[a][j][0], pa[a][0].x, ca[j][0].x);
                        acc[a][j][1] = sqdiff_acc(acc[a][j][1], pa[a][0].y, ca[j][0].y);
                        acc[a][j][2] = sqdiff_acc(acc[a][j][2], pa[a][0].z, ca[j][0].z);
                        acc[a][j][3] = sqdiff_acc(acc[a][j][3], pa[a][0].w, ca[j][0].w);
                        acc[a][j][4] = sqdiff_acc(acc[a][j][4], pa[a][1].x, ca[j][1].x);
                        acc[a][j][5] = sqdiff_acc(acc[a][j][5], pa[a][1].y, ca[j][1].y);
                        acc[a][j][6] = sqdiff_acc(acc[a][j][6], pa[a][1].z, ca[j][1].z);
                        acc[a][j][7] = sqdiff_acc(acc[a][j][7], pa[a][1].w, ca[j][1].w);
                    }
            }
        }
        float dist[2][4];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int j = 0; j < 4; j++) dist[a][j] = reduce8(acc[a][j]);
        // remainder dims: one 4-wide chunk (kmeans.rs:399-408) and a <4 tail (:411-416)
        __syncthreads();
        if (rem > 0) {
            load_rows(sp, kPT, npv, pt_row, data, D, Dmain, rem, false);
            load_rows(sc, kCT, ncv, c_row, cents, D, Dmain, rem, false);
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float r4 = 0.0f, tail = 0.0f;
                if (rem > 0) {
                    const float* p = &sp[(2 * tp + a) * kRowStride];
                    const float* c = &sc[(tc + 16 * j) * kRowStride];
                    int t0 = 0;
                    if (rem >= 4) {
                        float e0 = sqdiff_acc(0.0f, p[0], c[0]);
                        float e1 = sqdiff_acc(0.0f, p[1], c[1]);
                        float e2 = sqdiff_acc(0.0f, p[2], c[2]);
                        float e3 = sqdiff_acc(0.0f, p[3], c[3]);
                        r4 = __fadd_rn(__fadd_rn(__fadd_rn(e0, e1), e2), e3);
                        t0 = 4;
                    }
                    for (int t = t0; t < rem; t++) tail = sqdiff_acc(tail, p[t], c[t]);
                }
                dist[a][j] = __fadd_rn(__fadd_rn(dist[a][j], r4), tail);
            }
        if (MODE == 0) {
#pragma unroll
            for (int a = 0; a < 2; a++) {
                int pr = 2 * tp + a;
                if (pr < npv) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        int cr = tc + 16 * j;
                        if (cr < ncv) out[(size_t)(it.pt_off + p0 + pr) * ld + c0 + cr] = dist[a][j];
                    }
                }
            }
        } else {
#pragma unroll
            for (int a = 0; a < 2; a++) {
                int pr = 2 * tp + a;
                unsigned long long key = ~0ull;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int cr = tc + 16 * j;
                    if (cr < ncv) {
                        unsigned long long kk = ((unsigned long long)__float_as_uint(dist[a][j]) << 32) | (c0 + cr);
                        key = min(key, kk);
                    }
                }
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) key = min(key, __shfl_xor_sync(kFull, key, o));
                if (tc == 0 && pr < npv && key != ~0ull) {
                    uint32_t slot = it.pt_off + p0 + pr;
                    uint32_t pid = pt_entries ? pt_entries[slot].x : slot;
                    uint32_t hi_bits = pt_entries ? pt_entries[slot].y : 0u;
                    // NaN distances (bits > +inf) never win, as with the reference's '<'.
                    if ((key >> 32) < 0x7f800000ull) atomicMin(&best[pid], key | hi_bits);
                }
            }
        }
    }
}

// stable top-3 of each row of meta distances (kmeans.rs:651-672: stable sort, take 3)
__global__ void top3_kernel(const float* __restrict__ dist, uint64_t ld, uint32_t npts, uint32_t meta_k, uint32_t top,
                            uint32_t* __restrict__ out /* npts x 3 */) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    const float* row = dist + (size_t)i * ld;
    float d0 = 0, d1 = 0, d2 = 0;
    uint32_t i0 = kNoRow, i1 = kNoRow, i2 = kNoRow;
    for (uint32_t m = 0; m < meta_k; m++) {
        float d = row[m];
        // insert after entries with dist <= d (stable)
        if (i0 == kNoRow || d < d0) { d2 = d1; i2 = i1; d1 = d0; i1 = i0; d0 = d; i0 = m; }
        else if (i1 == kNoRow || d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = m; }
        else if (i2 == kNoRow || d < d2) { d2 = d; i2 = m; }
    }
    out[(size_t)i * 3 + 0] = i0;
    out[(size_t)i * 3 + 1] = top > 1 ? i1 : kNoRow;
    out[(size_t)i * 3 + 2] = top > 2 ? i2 : kNoRow;
}

// histogram of (point, rank) -> meta, then fill pt_entries grouped by meta
__global__ void meta_count_kernel(const uint32_t* __restrict__ top3, uint32_t npts, uint32_t* __restrict__ cnt) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts * 3u) return;
    uint32_t m = top3[i];
    if (m != kNoRow) atomicAdd(&cnt[m], 1u);
}
__global__ void meta_fill_kernel(const uint32_t* __restrict__ top3, uint32_t npts, uint32_t pt_base,
                                 const uint32_t* __restrict__ off, uint32_t* __restrict__ cur,
                                 uint2* __restrict__ entries) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts * 3u) return;
    uint32_t m = top3[i];
    if (m == kNoRow) return;
    uint32_t p = atomicAdd(&cur[m], 1u);
    entries[off[m] + p] = make_uint2(pt_base + i / 3u, (i % 3u) << 24);
}
// keys -> labels.  Flat mode (m2c == nullptr): label = low 32 bits.  Hierarchical mode:
// low bits = rank<<24 | j  ->  m2c_list[m2c_off[top3[p][rank]] + j]  (kmeans.rs:540-554).
__global__ void keys_to_labels_kernel(const unsigned long long* __restrict__ best, uint32_t npts,
                                      const uint32_t* __restrict__ top3, const uint32_t* __restrict__ m2c_off,
                                      const uint32_t* __restrict__ m2c_list, uint32_t* __restrict__ labels) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npts) return;
    uint32_t lowbits = (uint32_t)best[i];
    if (!m2c_off) { labels[i] = lowbits; return; }
    uint32_t rank = lowbits >> 24, j = lowbits & 0xffffffu;
    // kInitKey (no finite candidate): rank 0, j 0 = candidate_indices[0]
    uint32_t m = kNoRow;
    for (uint32_t r = rank; r < 3 && m == kNoRow; r++) {
        uint32_t mm = top3[(size_t)i * 3 + r];
        if (mm != kNoRow && m2c_off[mm + 1] > m2c_off[mm]) m = mm;
        if (r == rank && m == kNoRow) j = 0;
    }
    labels[i] = m == kNoRow ? 0u : m2c_list[m2c_off[m] + j];
}
__global__ void fill_u64_kernel(unsigned long long* p, unsigned long long v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// kmeans.rs:422-443: min_d[i] = min(min_d[i], dist_simd(data[i], latest))
__global__ void min_dist_kernel(const float* __restrict__ data, int D, uint32_t m, const float* __restrict__ latest,
                                float* __restrict__ min_d) {
    extern __shared__ float s_latest[];
    for (int d = threadIdx.x; d < D; d += blockDim.x) s_latest[d] = latest[d];
    __syncthreads();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float* p = data + (size_t)i * D;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int j = 0;
    bool al = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(data) & 15) == 0);
    if (al) {
        for (; j + 8 <= D; j += 8) {
            float4 x = __ldg(reinterpret_cast<const float4*>(p + j));
            float4 y = __ldg(reinterpret_cast<const float4*>(p + j + 4));
            a[0] = sqdiff_acc(a[0], x.x, s_latest[j + 0]);
            a[1] = sqdiff_acc(a[1], x.y, s_latest[j + 1]);
            a[2] = sqdiff_acc(a[2], x.z, s_latest[j + 2]);
            a[3] = sqdiff_acc(a[3], x.w, s_latest[j + 3]);
            a[4] = sqdiff_acc(a[4], y.x, s_latest[j + 4]);
            a[5] = sqdiff_acc(a[5], y.y, s_latest[j + 5]);
            a[6] = sqdiff_acc(a[6], y.z, s_latest[j + 6]);
            a[7] = sqdiff_acc(a[7], y.w, s_latest[j + 7]);
        }
    } else {
        for (; j + 8 <= D; j += 8)
#pragma unroll
            for (int l = 0; l < 8; l++) a[l] = sqdiff_acc(a[l], p[j + l], s_latest[j + l]);
    }
    float r8 = reduce8(a), r4 = 0.0f, tail = 0.0f;
    if (j + 4 <= D) {
        float e0 = sqdiff_acc(0.0f, p[j], s_latest[j]);
        float e1 = sqdiff_acc(0.0f, p[j + 1], s_latest[j + 1]);
        float e2 = sqdiff_acc(0.0f, p[j + 2], s_latest[j + 2]);
        float e3 = sqdiff_acc(0.0f, p[j + 3], s_latest[j + 3]);
        r4 = __fadd_rn(__fadd_rn(__fadd_rn(e0, e1), e2), e3);
        j += 4;
    }
    for (; j < D; j++) tail = sqdiff_acc(tail, p[j], s_latest[j]);
    float dist = __fadd_rn(__fadd_rn(r8, r4), tail);
    if (dist < min_d[i]) min_d[i] = dist;
}

// dst[dst_idx[i]] = src[src_idx[i]]   (rows of D floats); idx arrays may be nullptr (= i)
__global__ void copy_rows_kernel(const float* __restrict__ src, const uint32_t* __restrict__ src_idx,
                                 float* __restrict__ dst, const uint32_t* __restrict__ dst_idx, uint32_t nrows, int D) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nrows * D) return;
    uint32_t r = (uint32_t)(i / D);
    int d = (int)(i % D);
    size_t s = src_idx ? src_idx[r] : r, t = dst_idx ? dst_idx[r] : r;
    dst[t * D + d] = src[s * D + d];
}

// Per-cluster mean over a member list, members summed sequentially in list order
// (kmeans.rs:687-706 full batch, :749-776 mini-batch, :621-638 hierarchy).
// mode 0 (Lloyd):      out[c] = count ? sum / count : 0
// mode 1 (hierarchy):  out[c] = count ? sum / count : unchanged
// mode 2 (mini-batch): out[c] = (1-eta[c])*out[c] + eta[c]*(sum/count) for listed clusters
// cluster_ids: clusters to process (nullptr = 0..nclusters).
__global__ void cluster_mean_kernel(const float* __restrict__ data, int D, const uint32_t* __restrict__ member_off,
                                    const uint32_t* __restrict__ members, const uint32_t* __restrict__ cluster_ids,
                                    const float* __restrict__ eta, uint32_t nclusters, int mode, float* __restrict__ out) {
    uint32_t ci = blockIdx.x;
    if (ci >= nclusters) return;
    uint32_t c = cluster_ids ? cluster_ids[ci] : ci;
    uint32_t b = member_off[ci], e = member_off[ci + 1];
    uint32_t cnt = e - b;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = 0.0f;
        for (uint32_t t = b; t < e; t++) s = __fadd_rn(s, data[(size_t)members[t] * D + d]);
        if (mode == 0) {
            out[(size_t)c * D + d] = cnt ? __fdiv_rn(s, (float)cnt) : 0.0f;
        } else if (mode == 1) {
            if (cnt) out[(size_t)c * D + d] = __fdiv_rn(s, (float)cnt);
        } else if (cnt) {
            float mean = __fdiv_rn(s, (float)cnt);
            float et = eta[ci];
            float cur = out[(size_t)c * D + d];
            out[(size_t)c * D + d] = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, et), cur), __fmul_rn(et, mean));
        }
    }
}

// kmeans.rs:338-348 inner loop: per-centroid sum of squared movement (sequential in d).
__global__ void centroid_delta_kernel(const float* __restrict__ curr, const float* __restrict__ prev, uint32_t k, int D,
                                      float* __restrict__ local) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    float s = 0.0f;
    for (int d = 0; d < D; d++) s = sqdiff_acc(s, curr[(size_t)c * D + d], prev[(size_t)c * D + d]);
    local[c] = s;
}

// ====================================================================================
void launch_pairs(int mode, const float* data, int D, const float* cents, const PairItem* d_items,
                  const uint32_t* d_item_tile_off, int nitems, uint32_t total_tiles, const uint2* pt_entries,
                  const uint32_t* cent_idx, float* out, uint64_t ld, unsigned long long* best, cudaStream_t st) {
    if (!total_tiles) return;
    if (mode == 0)
        pairs_kernel<0><<<total_tiles, kPairsThreads, 0, st>>>(data, D, cents, d_items, d_item_tile_off, nitems, pt_entries,
                                                               cent_idx, out, ld, best);
    else
        pairs_kernel<1><<<total_tiles, kPairsThreads, 0, st>>>(data, D, cents, d_items, d_item_tile_off, nitems, pt_entries,
                                                               cent_idx, out, ld, best);
    VIDX_LAUNCHED();
}
uint32_t pairs_point_tile() { return kPT; }
void launch_top3(const float* dist, uint64_t ld, uint32_t npts, uint32_t meta_k, uint32_t top, uint32_t* out, cudaStream_t st) {
    if (!npts) return;
    top3_kernel<<<(unsigned)ceil_div(npts, 128), 128, 0, st>>>(dist, ld, npts, meta_k, top, out);
    VIDX_LAUNCHED();
}
void launch_meta_count(const uint32_t* top3, uint32_t npts, uint32_t* cnt, cudaStream_t st) {
    if (!npts) return;
    meta_count_kernel<<<(unsigned)ceil_div((size_t)npts * 3, 256), 256, 0, st>>>(top3, npts, cnt);
    VIDX_LAUNCHED();
}
void launch_meta_fill(const uint32_t* top3, uint32_t npts, uint32_t pt_base, const uint32_t* off, uint32_t* cur,
                      uint2* entries, cudaStream_t st) {
    if (!npts) return;
    meta_fill_kernel<<<(unsigned)ceil_div((size_t)npts * 3, 256), 256, 0, st>>>(top3, npts, pt_base, off, cur, entries);
    VIDX_LAUNCHED();
}
void launch_keys_to_labels(const unsigned long long* best, uint32_t npts, const uint32_t* top3, const uint32_t* m2c_off,
                           const uint32_t* m2c_list, uint32_t* labels, cudaStream_t st) {
    if (!npts) return;
    keys_to_labels_kernel<<<(unsigned)ceil_div(npts, 256), 256, 0, st>>>(best, npts, top3, m2c_off, m2c_list, labels);
    VIDX_LAUNCHED();
}
void launch_fill_keys(unsigned long long* p, size_t n, cudaStream_t st) {
    if (!n) return;
    fill_u64_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(p, kInitKey, n);
    VIDX_LAUNCHED();
}
void launch_min_dist(const float* data, int D, uint32_t m, const float* latest, float* min_d, cudaStream_t st) {
    if (!m) return;
    min_dist_kernel<<<(unsigned)ceil_div(m, 128), 128, (size_t)D * 4, st>>>(data, D, m, latest, min_d);
    VIDX_LAUNCHED();
}
void launch_copy_rows(const float* src, const uint32_t* src_idx, float* dst, const uint32_t* dst_idx, uint32_t nrows, int D,
                      cudaStream_t st) {
    if (!nrows) return;
    copy_rows_kernel<<<(unsigned)ceil_div((size_t)nrows * D, 256), 256, 0, st>>>(src, src_idx, dst, dst_idx, nrows, D);
    VIDX_LAUNCHED();
}
void launch_cluster_mean(const float* data, int D, const uint32_t* member_off, const uint32_t* members,
                         const uint32_t* cluster_ids, const float* eta, uint32_t nclusters, int mode, float* out,
                         cudaStream_t st) {
    if (!nclusters) return;
    int threads = D >= 128 ? 128 : (D >= 64 ? 64 : 32);
    cluster_mean_kernel<<<nclusters, threads, 0, st>>>(data, D, member_off, members, cluster_ids, eta, nclusters, mode, out);
    VIDX_LAUNCHED();
}
void launch_centroid_delta(const float* curr, const float* prev, uint32_t k, int D, float* local, cudaStream_t st) {
    if (!k) return;
    centroid_delta_kernel<<<(unsigned)ceil_div(k, 128), 128, 0, st>>>(curr, prev, k, D, local);
    VIDX_LAUNCHED();
}

}  // namespace vidx

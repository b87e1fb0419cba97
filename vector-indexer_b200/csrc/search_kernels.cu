// search_kernels.cu -- IVF search on sm_100a: coarse quantization, probe selection,
// (query,segment) grouping, the inverted-list scan with fused in-register top-k, and
// the per-query merge.  Replaces IvfIndex::search_with_paths (src/ivf_index.rs:190-267)
// and euclidean_distance_squared (src/utils.rs:28-30) for a whole batch of queries.
//
// Arithmetic contract: every distance is the reference's strictly sequential left fold
//   acc = acc + (x - y)*(x - y)            (3 roundings per element, no FMA)
// in dimension order, so distances are bit-identical to the reference and every
// ordering decision (probe order, top-k order) is made on identical keys.
//
// Data layout (HBM): see f4_index() in common.cuh.  Vectors of one list are stored in
// consecutive supergroups of 128 vectors, each float4 [Dq][128]; a "group" is 32 of those
// vectors (one per lane), so a warp's 16-byte loads for one chunk are 512 contiguous bytes
// whether they come from HBM directly (sparse kernel) or from the shared-memory stage (dense
// kernel).  Lists are padded to whole supergroups (zero rows, masked by position); a "segment"
// is up to 32 groups (1024 vectors) of one list and is the unit of work and of result slots.
#include "search.h"

namespace vidx {

// ------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------
__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// xq [nq][D] -> xqp [nq][Dp] zero padded (only needed when D % 4 != 0).
__global__ void pad_rows_kernel(const float* __restrict__ in, float* __restrict__ out, uint64_t nrows, int D, int Dp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * (size_t)Dp) return;
    size_t r = i / Dp;
    int d = (int)(i % Dp);
    out[i] = d < D ? in[r * D + d] : 0.0f;
}

// ---- device-wide exclusive scan of u32 (3 kernels; out has n+1 entries) -------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* warp_sums) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(kFull, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(kFull, s, o);
            if (lane >= o) s += y;
        }
        warp_sums[lane] = s;  // inclusive
    }
    __syncthreads();
    uint32_t base = w ? warp_sums[w - 1] : 0;
    if (total) *total = warp_sums[(blockDim.x >> 5) - 1];
    uint32_t r = base + x - v;
    __syncthreads();
    return r;
}

__global__ void scan_block_sums_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ bsum, size_t n) {
    __shared__ uint32_t ws[32];
    size_t base = (size_t)blockIdx.x * kScanTile;
    uint32_t s = 0;
    for (int i = 0; i < kScanItems; i++) {
        size_t j = base + (size_t)i * kScanThreads + threadIdx.x;
        if (j < n) s += in[j];
    }
    uint32_t tot;
    block_exclusive_scan(s, &tot, ws);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
// single block: exclusive scan of bsum[nb] in place (+ total at bsum[nb])
__global__ void scan_of_sums_kernel(uint32_t* bsum, size_t nb) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < nb; base += blockDim.x) {
        size_t j = base + threadIdx.x;
        uint32_t v = j < nb ? bsum[j] : 0;
        uint32_t tot;
        uint32_t e = block_exclusive_scan(v, &tot, ws);
        uint32_t c = carry;
        if (j < nb) bsum[j] = c + e;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[nb] = carry;
}
__global__ void scan_apply_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ bsum,
                                  uint32_t* __restrict__ out, size_t n) {
    __shared__ uint32_t ws[32];
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    uint32_t e = block_exclusive_scan(s, nullptr, ws) + bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) out[base + i] = e;
        e += v[i];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = bsum[gridDim.x];
}

// n <= kScanTile: one block, one launch (the per-list prefix sums of a search are this short)
__global__ void scan_single_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
    __shared__ uint32_t ws[32];
    size_t base = (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    uint32_t tot;
    uint32_t e = block_exclusive_scan(s, &tot, ws);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) out[base + i] = e;
        e += v[i];
    }
    if (threadIdx.x == 0) out[n] = tot;
}

void exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, size_t n, uint32_t* d_tmp, cudaStream_t st) {
    if (n <= (size_t)kScanTile) {
        scan_single_kernel<<<1, kScanThreads, 0, st>>>(d_in, d_out, n);
        VIDX_LAUNCHED();
        return;
    }
    size_t nb = ceil_div(n, kScanTile);
    if (nb == 0) nb = 1;
    scan_block_sums_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(d_in, d_tmp, n);
    VIDX_LAUNCHED();
    scan_of_sums_kernel<<<1, kScanThreads, 0, st>>>(d_tmp, nb);
    VIDX_LAUNCHED();
    scan_apply_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(d_in, d_tmp, d_out, n);
    VIDX_LAUNCHED();
}
size_t exclusive_scan_tmp_entries(size_t n) { return ceil_div(n, kScanTile) + 2; }

// ------------------------------------------------------------------------------------
// tile machinery shared by the coarse-distance kernel and the dense scan kernel
// ------------------------------------------------------------------------------------
// A block of 8 warps works on a tile of 4 groups (128 vectors) x up to 64 queries.
// Warp w owns queries 8w..8w+7 and all 128 vectors: lane l holds the 4x8 accumulators
// of vectors {32 i + l}.  The dimension axis is streamed through shared memory in
// chunks of 8 float4 (32 dims) with a 3-deep cp.async ring.
constexpr int kTileGroups = 4;
constexpr int kChunkF4 = 8;                               // float4 per vector per stage
constexpr int kTileQ = 64;
constexpr int kStages = 3;
constexpr int kVecStageF4 = kTileGroups * kChunkF4 * 32;  // 1024 float4 = 16 KB
constexpr int kQStageF4 = kTileQ * kChunkF4;              // 512 float4  =  8 KB
constexpr int kStageF4 = kVecStageF4 + kQStageF4;
constexpr int kDenseSmemBytes = kStages * kStageF4 * 16;  // 72 KB
constexpr int kDenseThreads = 256;

// Issue the copies of one stage: the 4 groups' float4 [c0, c0+8) and the tile's queries.
// ng_tile: groups present in this tile (<=4); nqt: queries present (<=64).
__device__ __forceinline__ void stage_issue(float4* stage, const float4* __restrict__ vecs, size_t gbase, int ng_tile,
                                            int Dq, int c0, const float4* __restrict__ xq4, const uint32_t* q_ids,
                                            int nqt) {
    float4* sv = stage;
    float4* sq = stage + kVecStageF4;
    int tid = threadIdx.x;
#pragma unroll
    for (int it = 0; it < kVecStageF4 / kDenseThreads; it++) {
        int idx = it * kDenseThreads + tid;
        int i = idx >> 8;          // group within tile (256 float4 per group per stage)
        int c = (idx >> 5) & 7;    // float4 within chunk
        int l = idx & 31;
        if (i < ng_tile && c0 + c < Dq) cp_async16(&sv[idx], &vecs[f4_index(gbase + i, Dq, c0 + c, l)]);
    }
#pragma unroll
    for (int it = 0; it < kQStageF4 / kDenseThreads; it++) {
        int idx = it * kDenseThreads + tid;
        int qi = idx >> 3, c = idx & 7;
        if (qi < nqt && c0 + c < Dq) cp_async16(&sq[idx], &xq4[(size_t)q_ids[qi] * Dq + c0 + c]);
    }
}

template <int QT>
__device__ __forceinline__ void tile_compute(const float4* stage, int warp, int lane, int nc, float (&acc)[4][8]) {
    const float4* sv = stage;
    const float4* sq = stage + kVecStageF4 + warp * 8 * kChunkF4;
#pragma unroll 2
    for (int c = 0; c < nc; c++) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = sv[(i * kChunkF4 + c) * 32 + lane];
#pragma unroll
        for (int j = 0; j < QT; j++) {
            float4 q = sq[j * kChunkF4 + c];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float a = acc[i][j];
                a = sqdiff_acc(a, q.x, v[i].x);
                a = sqdiff_acc(a, q.y, v[i].y);
                a = sqdiff_acc(a, q.z, v[i].z);
                a = sqdiff_acc(a, q.w, v[i].w);
                acc[i][j] = a;
            }
        }
    }
}

__device__ __forceinline__ void tile_compute_dispatch(int nqw, const float4* stage, int warp, int lane, int nc,
                                                      float (&acc)[4][8]) {
    switch (nqw) {
        case 8: tile_compute<8>(stage, warp, lane, nc, acc); break;
        case 7: tile_compute<7>(stage, warp, lane, nc, acc); break;
        case 6: tile_compute<6>(stage, warp, lane, nc, acc); break;
        case 5: tile_compute<5>(stage, warp, lane, nc, acc); break;
        case 4: tile_compute<4>(stage, warp, lane, nc, acc); break;
        case 3: tile_compute<3>(stage, warp, lane, nc, acc); break;
        case 2: tile_compute<2>(stage, warp, lane, nc, acc); break;
        case 1: tile_compute<1>(stage, warp, lane, nc, acc); break;
        default: break;
    }
}

// ------------------------------------------------------------------------------------
// K1: coarse quantization -- exact distances of every query to every centroid
// (src/ivf_index.rs:205-213).  Centroids are stored in the same interleaved group
// layout.  grid = (ceil(ngroups/4), ceil(nq/64)); out[q][c], row stride ldo.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDenseThreads, 2)
coarse_dist_kernel(const float4* __restrict__ cents, int ngroups, int Dq, const float4* __restrict__ xq4, uint32_t nq,
                   float* __restrict__ out, uint32_t ldo) {
    extern __shared__ __align__(16) float4 smem[];
    __shared__ uint32_t q_ids[kTileQ];
    int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    size_t gbase = (size_t)blockIdx.x * kTileGroups;
    int ng_tile = min(kTileGroups, ngroups - (int)gbase);
    uint32_t q0 = blockIdx.y * kTileQ;
    int nqt = (int)min((uint32_t)kTileQ, nq - q0);
    if (tid < kTileQ) q_ids[tid] = q0 + tid;
    __syncthreads();
    int nqw = max(0, min(8, nqt - warp * 8));
    int ndc = (Dq + kChunkF4 - 1) / kChunkF4;

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;

    stage_issue(smem, cents, gbase, ng_tile, Dq, 0, xq4, q_ids, nqt);
    cp_async_commit();
    if (ndc > 1) stage_issue(smem + kStageF4, cents, gbase, ng_tile, Dq, kChunkF4, xq4, q_ids, nqt);
    cp_async_commit();
    for (int s = 0; s < ndc; s++) {
        cp_async_wait<1>();
        __syncthreads();
        if (s + 2 < ndc)
            stage_issue(smem + ((s + 2) % kStages) * kStageF4, cents, gbase, ng_tile, Dq, (s + 2) * kChunkF4, xq4, q_ids,
                        nqt);
        cp_async_commit();
        int nc = min(kChunkF4, Dq - s * kChunkF4);
        tile_compute_dispatch(nqw, smem + (s % kStages) * kStageF4, warp, lane, nc, acc);
    }
    cp_async_wait<0>();
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < nqw) {
            size_t row = (size_t)(q0 + warp * 8 + j) * ldo;
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (i < ng_tile) out[row + (gbase + i) * 32 + lane] = acc[i][j];
        }
    }
}

// ------------------------------------------------------------------------------------
// K2 / generic exact top-k of a row of non-negative floats by (value, position):
// the stable ascending sort + take(k) of src/ivf_index.rs:215-220 and :265-266.
// One block per row.  Keys are u64 (float bits << 32 | position), all distinct, so an
// MSB-first radix select finds the k-th smallest key exactly; the <= k survivors are
// then bitonic-sorted in shared memory.
// ------------------------------------------------------------------------------------
constexpr int kSelThreads = 256;

__device__ __forceinline__ uint64_t make_key(float d, uint32_t pos) {
    return ((uint64_t)__float_as_uint(d) << 32) | pos;
}

// rows: row r starts at vals + row_off[r] (or r*ld when row_off == nullptr) and has
// row_len[r] (or n) entries.  Entries with NaN or +inf value are never selected.
// out_pos[r*k + t], out_val[r*k + t]: t-th smallest; padded with kNoRow / +inf.
__global__ void __launch_bounds__(kSelThreads)
select_topk_kernel(const float* __restrict__ vals, const uint64_t* __restrict__ row_off, const uint32_t* __restrict__ row_len,
                   uint64_t ld, uint32_t n_fixed, uint32_t k, uint32_t kcap /*pow2 >= k*/, uint32_t* __restrict__ out_pos,
                   float* __restrict__ out_val) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sel_smem);  // kcap entries
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_count;
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_need;

    size_t r = blockIdx.x;
    const float* row = vals + (row_off ? row_off[r] : r * ld);
    uint32_t n = row_len ? row_len[r] : n_fixed;
    int tid = threadIdx.x;
    const uint64_t kInfKey = (uint64_t)0x7f800000u << 32;  // keys >= this are +inf / NaN

    // Number of selectable (finite) entries.
    uint32_t local = 0;
    for (uint32_t i = tid; i < n; i += kSelThreads) local += (make_key(row[i], 0) < kInfKey) ? 1u : 0u;
    if (tid == 0) s_count = 0;
    __syncthreads();
    atomicAdd(&s_count, local);
    __syncthreads();
    uint32_t nfinite = s_count;
    uint32_t kk = min(k, nfinite);
    __syncthreads();

    uint64_t thresh;  // select keys <= thresh
    if (kk == nfinite) {
        thresh = kInfKey - 1;
    } else {
        // radix select the kk-th smallest key (1-based rank kk), 8 bits at a time.
        if (tid == 0) { s_prefix = 0; s_need = kk; }
        __syncthreads();
        for (int shift = 56; shift >= 0; shift -= 8) {
            hist[tid] = 0;
            __syncthreads();
            uint64_t prefix = s_prefix;
            uint64_t himask = shift == 56 ? 0ull : (~0ull << (shift + 8));
            for (uint32_t i = tid; i < n; i += kSelThreads) {
                uint64_t key = make_key(row[i], i);
                if (key < kInfKey && (key & himask) == prefix) atomicAdd(&hist[(key >> shift) & 0xff], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t need = s_need, cum = 0;
                int b = 0;
                for (; b < 256; b++) {
                    if (cum + hist[b] >= need) break;
                    cum += hist[b];
                }
                s_need = need - cum;
                s_prefix = prefix | ((uint64_t)b << shift);
            }
            __syncthreads();
            // Early exit: if the chosen digit bucket holds exactly the remaining need, all
            // keys with this prefix are selected.
            if (shift > 0) {
                uint64_t p2 = s_prefix;
                uint32_t b = (uint32_t)((p2 >> shift) & 0xff);
                if (hist[b] == s_need) {
                    __syncthreads();
                    if (tid == 0) s_prefix = p2 | ((1ull << shift) - 1);
                    __syncthreads();
                    break;
                }
            }
            __syncthreads();
        }
        thresh = s_prefix;
    }
    __syncthreads();
    // collect
    if (tid == 0) s_count = 0;
    for (uint32_t i = tid; i < kcap; i += kSelThreads) keys[i] = ~0ull;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += kSelThreads) {
        uint64_t key = make_key(row[i], i);
        if (key <= thresh && key < kInfKey) {
            uint32_t p = atomicAdd(&s_count, 1u);
            if (p < kcap) keys[p] = key;
        }
    }
    __syncthreads();
    // bitonic sort of kcap keys
    for (uint32_t size = 2; size <= kcap; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = tid; i < kcap / 2; i += kSelThreads) {
                uint32_t lo = 2 * i - (i & (stride - 1));
                uint32_t hi = lo + stride;
                bool up = (lo & size) == 0;
                uint64_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (uint32_t t = tid; t < k; t += kSelThreads) {
        uint64_t key = t < kk ? keys[t] : ~0ull;
        bool ok = key != ~0ull;
        out_pos[r * k + t] = ok ? (uint32_t)key : kNoRow;
        if (out_val) out_val[r * k + t] = ok ? __uint_as_float((uint32_t)(key >> 32)) : __int_as_float(0x7f800000);
    }
}

// Short rows, k <= 32: one warp per row keeps the 32 smallest (value, position) pairs sorted across its lanes (lane t =
// t-th smallest) and inserts a candidate with one ballot and one shuffle; no shared memory, no block barriers.  Same
// contract as select_topk_kernel (ties by position, +inf / NaN never selected, padding kNoRow / +inf); values of either
// sign.
__global__ void __launch_bounds__(256)
select_small_kernel(const float* __restrict__ vals, const uint64_t* __restrict__ row_off, const uint32_t* __restrict__ row_len,
                    uint64_t ld, uint32_t n_fixed, uint64_t nrows, uint32_t k, uint32_t* __restrict__ out_pos, float* __restrict__ out_val) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= nrows) return;
    const float* row = vals + (row_off ? row_off[r] : r * ld);
    const uint32_t n = row_len ? row_len[r] : n_fixed;
    const float kInf = __int_as_float(0x7f800000);
    float my_d = kInf;
    uint32_t my_p = kNoRow;
    auto less = [](float d, uint32_t p, float d2, uint32_t p2) { return d < d2 || (d == d2 && p < p2); };
    uint32_t first = 0;
    if (n) {
        // the first 32 entries: a bitonic sort across the lanes instead of 32 serial insertions
        const float v = (uint32_t)lane < n ? row[lane] : kInf;
        if (v < kInf) {
            my_d = v;
            my_p = (uint32_t)lane;
        }
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const float od = __shfl_xor_sync(kFull, my_d, stride);
                const uint32_t op = __shfl_xor_sync(kFull, my_p, stride);
                const bool up = (lane & size) == 0;            // ascending block
                const bool lower = (lane & stride) == 0;       // this lane keeps the smaller of the pair in an ascending block
                const bool take_min = up == lower;
                const bool o_less = less(od, op, my_d, my_p);
                if (take_min == o_less) {
                    my_d = od;
                    my_p = op;
                }
            }
        }
        first = 32;
    }
    // four loads in flight per lane: with few rows (small batches) one warp's chain of dependent loads is the run time
    for (uint32_t base4 = first; base4 < n; base4 += 128) {
      float v4[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
          const uint32_t i = base4 + 32u * u + lane;
          v4[u] = i < n ? row[i] : kInf;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t base = base4 + 32u * u;
        if (base >= n) break;
        const uint32_t i = base + lane;
        const float v = v4[u];
        float kd = __shfl_sync(kFull, my_d, (int)k - 1);
        uint32_t kp = __shfl_sync(kFull, my_p, (int)k - 1);
        unsigned m = __ballot_sync(kFull, v < kInf && less(v, i, kd, kp));
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cd = __shfl_sync(kFull, v, src);
            const uint32_t cp = base + (uint32_t)src;
            const unsigned gm = __ballot_sync(kFull, less(cd, cp, my_d, my_p));  // lanes whose pair is greater: a suffix
            if (!gm) continue;
            const int pos = __ffs(gm) - 1;
            const float up_d = __shfl_up_sync(kFull, my_d, 1);
            const uint32_t up_p = __shfl_up_sync(kFull, my_p, 1);
            if (lane == pos) {
                my_d = cd;
                my_p = cp;
            } else if (lane > pos) {
                my_d = up_d;
                my_p = up_p;
            }
        }
      }
    }
    if (lane < (int)k) {
        if (out_pos) out_pos[r * k + lane] = my_p;
        if (out_val) out_val[r * k + lane] = my_d;
    }
}

// ------------------------------------------------------------------------------------
// K3: grouping (query, probed list) -> per-segment query lists + result slots
// (src/ivf_index.rs:223-246 groups probes by shard; here the unit is the segment).
// ------------------------------------------------------------------------------------
// pair p = q*nprobe + r.  pair_ns[p] = segments of the probed list (0 for padding / lists
// this rank does not own).  seg_cnt[s] += 1 for each of them.
// only_flag (optional): per-query flag; when given, only queries whose flag is set take part
// (the queries the tensor-core path handed back to the exact kernels).
__global__ void group_count_kernel(const uint32_t* __restrict__ probes, size_t npairs, uint32_t nprobe,
                                   const uint2* __restrict__ list_seg, const uint32_t* __restrict__ only_flag,
                                   uint32_t* __restrict__ pair_ns, uint32_t* __restrict__ seg_cnt) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    uint32_t l = probes[p];
    uint32_t ns = 0;
    if (l != kNoRow && (!only_flag || only_flag[p / nprobe])) {
        uint2 sr = list_seg[l];
        uint32_t s0 = sr.x, s1 = sr.y;
        ns = s1 - s0;
        for (uint32_t s = s0; s < s1; s++) atomicAdd(&seg_cnt[s], 1u);
    }
    pair_ns[p] = ns;
}
// seg_qlist[seg_qoff[s] + i] = (query, slot).  slot = slot_off[p] + (s - s0).
__global__ void group_fill_kernel(const uint32_t* __restrict__ probes, size_t npairs, uint32_t nprobe,
                                  const uint2* __restrict__ list_seg, const uint32_t* __restrict__ slot_off,
                                  const uint32_t* __restrict__ seg_qoff, uint32_t* __restrict__ seg_cur,
                                  uint2* __restrict__ seg_qlist, uint32_t* __restrict__ slot_seg,
                                  const uint32_t* __restrict__ only_flag, uint32_t* __restrict__ slot_rank) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    uint32_t l = probes[p];
    if (l == kNoRow) return;
    uint32_t q = (uint32_t)(p / nprobe);
    if (only_flag && !only_flag[q]) return;
    uint2 sr = list_seg[l];
    uint32_t s0 = sr.x, s1 = sr.y;
    uint32_t slot = slot_off[p];
    for (uint32_t s = s0; s < s1; s++, slot++) {
        uint32_t i = atomicAdd(&seg_cur[s], 1u);
        seg_qlist[seg_qoff[s] + i] = make_uint2(q, slot);
        slot_seg[slot] = s;
        slot_rank[slot] = (uint32_t)(p - (size_t)q * nprobe);
    }
}
// One thread per segment: split its query list into dense items (tiles of <= 64 queries
// handled by one block) and sparse items (<= 8 queries, the block's warps split the
// vectors instead).  counters[0] = #dense, counters[1] = #sparse.
__global__ void group_items_kernel(const uint32_t* __restrict__ seg_cnt, uint32_t nseg, uint32_t sparse_max,
                                   ScanItem* __restrict__ dense, ScanItem* __restrict__ sparse,
                                   uint32_t* __restrict__ counters) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    uint32_t c = seg_cnt[s];
    if (c == 0) return;
    uint32_t nfull = c / kTileQ, rem = c % kTileQ;
    uint32_t nd = nfull + ((rem > sparse_max) ? 1u : 0u);
    uint32_t nsp = (rem > 0 && rem <= sparse_max) ? (rem + 7) / 8 : 0u;
    if (nd) {
        uint32_t b = atomicAdd(&counters[0], nd);
        for (uint32_t t = 0; t < nd; t++) {
            ScanItem it;
            it.seg = s;
            it.qstart = t * kTileQ;
            it.nq = min((uint32_t)kTileQ, c - t * kTileQ);
            it.pad = 0;
            dense[b + t] = it;
        }
    }
    if (nsp) {
        uint32_t b = atomicAdd(&counters[1], nsp);
        for (uint32_t t = 0; t < nsp; t++) {
            ScanItem it;
            it.seg = s;
            it.qstart = nfull * kTileQ + t * 8;
            it.nq = min(8u, c - it.qstart);
            it.pad = 0;
            sparse[b + t] = it;
        }
    }
}

// ------------------------------------------------------------------------------------
// K4+K5 dense: list scan with fused top-k for segments probed by many queries.
// Persistent blocks pull items {segment, tile of <= 64 of its queries}.  Replaces the
// scan loop + sort of src/ivf_index.rs:252-266 (distance: src/utils.rs:28-30).
// ------------------------------------------------------------------------------------
// Selection of one tile's 4x(QT) register distances into the warp lists.
template <bool ALLDIST>
__device__ __forceinline__ void tile_select(float (&acc)[4][8], int nqw, int lane, uint32_t rowbase, uint32_t posbase,
                                            uint32_t nvalid, uint32_t k, float (&my_d)[8], uint32_t (&my_r)[8],
                                            float (&thr)[8], float* (&alld)[8]) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < nqw) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t pos = posbase + i * 32 + lane;
                bool valid = pos < nvalid;
                float d = acc[i][j];
                if (ALLDIST) {
                    if (pos < (uint32_t)kSegVecs) alld[j][pos] = valid ? d : __int_as_float(0x7f800000);
                } else {
                    unsigned m = __ballot_sync(kFull, valid && d < thr[j]);
                    while (m) {
                        int src = __ffs(m) - 1;
                        m &= m - 1;
                        float cd = __shfl_sync(kFull, d, src);
                        if (cd < thr[j]) {
                            warp_insert_stable(cd, rowbase + i * 32 + src, my_d[j], my_r[j], lane);
                            thr[j] = __shfl_sync(kFull, my_d[j], k - 1);
                        }
                    }
                }
            }
        }
    }
}

template <bool ALLDIST>
__global__ void __launch_bounds__(kDenseThreads, 2)
scan_dense_kernel(const float4* __restrict__ vecs, int Dq, const float4* __restrict__ xq4,
                  const SegDesc* __restrict__ segs, const uint32_t* __restrict__ seg_qoff,
                  const uint2* __restrict__ seg_qlist, const ScanItem* __restrict__ items,
                  const uint32_t* __restrict__ counters, uint32_t* __restrict__ work_counter, uint32_t k,
                  float* __restrict__ cand_d, uint32_t* __restrict__ cand_r, float* __restrict__ alldist) {
    extern __shared__ __align__(16) float4 smem[];
    __shared__ uint32_t q_ids[kTileQ];
    __shared__ uint32_t q_slot[kTileQ];
    __shared__ uint32_t s_item;
    int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t n_items = counters[0];
    int ndc = (Dq + kChunkF4 - 1) / kChunkF4;

    for (;;) {
        if (tid == 0) s_item = atomicAdd(work_counter, 1u);
        __syncthreads();
        uint32_t item = s_item;
        if (item >= n_items) break;
        ScanItem it = items[item];
        SegDesc sg = segs[it.seg];
        int nqt = (int)it.nq;
        if (tid < nqt) {
            uint2 e = seg_qlist[seg_qoff[it.seg] + it.qstart + tid];
            q_ids[tid] = e.x;
            q_slot[tid] = e.y;
        }
        __syncthreads();
        int nqw = max(0, min(8, nqt - warp * 8));
        float my_d[8], thr[8];
        uint32_t my_r[8];
        float* alld[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            my_d[j] = __int_as_float(0x7f800000);
            thr[j] = __int_as_float(0x7f800000);
            my_r[j] = kNoRow;
            alld[j] = nullptr;
            if (ALLDIST && j < nqw) alld[j] = alldist + (size_t)q_slot[warp * 8 + j] * kSegVecs;
        }
        int ntiles = ((int)sg.ng + kTileGroups - 1) / kTileGroups;
        int total = ntiles * ndc;
        float acc[4][8];

        {
            int ngt = min(kTileGroups, (int)sg.ng);
            stage_issue(smem, vecs, sg.g0, ngt, Dq, 0, xq4, q_ids, nqt);
            cp_async_commit();
            if (total > 1) {
                int t1 = 1 / ndc, d1 = 1 % ndc;
                stage_issue(smem + kStageF4, vecs, (size_t)sg.g0 + t1 * kTileGroups,
                            min(kTileGroups, (int)sg.ng - t1 * kTileGroups), Dq, d1 * kChunkF4, xq4, q_ids, nqt);
            }
            cp_async_commit();
        }
        int t = 0, dc = 0;
        for (int s = 0; s < total; s++) {
            cp_async_wait<1>();
            __syncthreads();
            if (s + 2 < total) {
                int t2 = (s + 2) / ndc, d2 = (s + 2) % ndc;
                stage_issue(smem + ((s + 2) % kStages) * kStageF4, vecs, (size_t)sg.g0 + t2 * kTileGroups,
                            min(kTileGroups, (int)sg.ng - t2 * kTileGroups), Dq, d2 * kChunkF4, xq4, q_ids, nqt);
            }
            cp_async_commit();
            if (dc == 0) {
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;
            }
            int nc = min(kChunkF4, Dq - dc * kChunkF4);
            tile_compute_dispatch(nqw, smem + (s % kStages) * kStageF4, warp, lane, nc, acc);
            if (dc == ndc - 1) {
                uint32_t posbase = (uint32_t)t * kTileGroups * 32;
                tile_select<ALLDIST>(acc, nqw, lane, (sg.g0 + t * kTileGroups) * 32u, posbase, sg.nvalid, k, my_d, my_r,
                                     thr, alld);
            }
            if (++dc == ndc) { dc = 0; t++; }
        }
        cp_async_wait<0>();
        if (!ALLDIST) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j < nqw && lane < (int)k) {
                    size_t o = (size_t)q_slot[warp * 8 + j] * k + lane;
                    cand_d[o] = my_d[j];
                    cand_r[o] = my_r[j];
                }
            }
        } else {
            // positions beyond the last tile of a short segment
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < nqw)
                    for (uint32_t pos = (uint32_t)ntiles * kTileGroups * 32 + lane; pos < (uint32_t)kSegVecs; pos += 32)
                        alld[j][pos] = __int_as_float(0x7f800000);
        }
        __syncthreads();  // q_ids / smem reuse
    }
}

// ------------------------------------------------------------------------------------
// K4+K5 sparse: segments probed by few (<= 8) queries per item.  The work is HBM-bound:
// the block's 8 warps split the segment's groups and stream them straight from HBM with
// coalesced 16-byte loads (512 B per warp instruction); queries sit in shared memory.
// Warp lists are merged across the block at the end (ordered by (dist, row)).
// ------------------------------------------------------------------------------------
constexpr int kSparseThreads = 256;

template <int QT>
__device__ __forceinline__ void sparse_group_pair(const float4* __restrict__ v0, const float4* __restrict__ v1, bool has1,
                                                  int Dq, const float4* sq, int lane, float (&a0)[8], float (&a1)[8]) {
#pragma unroll
    for (int j = 0; j < 8; j++) { a0[j] = 0.0f; a1[j] = 0.0f; }
    int c = 0;
    for (; c + 4 <= Dq; c += 4) {
        float4 x0[4], x1[4];
#pragma unroll
        for (int u = 0; u < 4; u++) x0[u] = ldg_f4(&v0[(size_t)(c + u) * kSuper + lane]);
#pragma unroll
        for (int u = 0; u < 4; u++) x1[u] = has1 ? ldg_f4(&v1[(size_t)(c + u) * kSuper + lane]) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int j = 0; j < QT; j++) {
                float4 q = sq[(size_t)j * Dq + c + u];
                float a = a0[j];
                a = sqdiff_acc(a, q.x, x0[u].x);
                a = sqdiff_acc(a, q.y, x0[u].y);
                a = sqdiff_acc(a, q.z, x0[u].z);
                a = sqdiff_acc(a, q.w, x0[u].w);
                a0[j] = a;
                float b = a1[j];
                b = sqdiff_acc(b, q.x, x1[u].x);
                b = sqdiff_acc(b, q.y, x1[u].y);
                b = sqdiff_acc(b, q.z, x1[u].z);
                b = sqdiff_acc(b, q.w, x1[u].w);
                a1[j] = b;
            }
        }
    }
    for (; c < Dq; c++) {
        float4 x0 = ldg_f4(&v0[(size_t)c * kSuper + lane]);
        float4 x1 = has1 ? ldg_f4(&v1[(size_t)c * kSuper + lane]) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < QT; j++) {
            float4 q = sq[(size_t)j * Dq + c];
            float a = a0[j];
            a = sqdiff_acc(a, q.x, x0.x);
            a = sqdiff_acc(a, q.y, x0.y);
            a = sqdiff_acc(a, q.z, x0.z);
            a = sqdiff_acc(a, q.w, x0.w);
            a0[j] = a;
            float b = a1[j];
            b = sqdiff_acc(b, q.x, x1.x);
            b = sqdiff_acc(b, q.y, x1.y);
            b = sqdiff_acc(b, q.z, x1.z);
            b = sqdiff_acc(b, q.w, x1.w);
            a1[j] = b;
        }
    }
}

template <bool ALLDIST>
__device__ __forceinline__ void sparse_select(float (&a)[8], int nq, int lane, uint32_t rowbase, uint32_t posbase,
                                              uint32_t nvalid, uint32_t k, float (&my_d)[8], uint32_t (&my_r)[8],
                                              float (&thr)[8], float* (&alld)[8]) {
    uint32_t pos = posbase + lane;
    bool valid = pos < nvalid;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < nq) {
            float d = a[j];
            if (ALLDIST) {
                alld[j][pos] = valid ? d : __int_as_float(0x7f800000);
            } else {
                unsigned m = __ballot_sync(kFull, valid && d < thr[j]);
                while (m) {
                    int src = __ffs(m) - 1;
                    m &= m - 1;
                    float cd = __shfl_sync(kFull, d, src);
                    if (cd < thr[j]) {
                        warp_insert_stable(cd, rowbase + src, my_d[j], my_r[j], lane);
                        thr[j] = __shfl_sync(kFull, my_d[j], k - 1);
                    }
                }
            }
        }
    }
}

template <bool ALLDIST>
__global__ void __launch_bounds__(kSparseThreads, 2)
scan_sparse_kernel(const float4* __restrict__ vecs, int Dq, const float4* __restrict__ xq4,
                   const SegDesc* __restrict__ segs, const uint32_t* __restrict__ seg_qoff,
                   const uint2* __restrict__ seg_qlist, const ScanItem* __restrict__ items,
                   const uint32_t* __restrict__ counters, uint32_t* __restrict__ work_counter, uint32_t k,
                   float* __restrict__ cand_d, uint32_t* __restrict__ cand_r, float* __restrict__ alldist) {
    extern __shared__ __align__(16) float4 smem[];  // queries: 8 x Dq float4, then merge area
    __shared__ uint32_t q_ids[8];
    __shared__ uint32_t q_slot[8];
    __shared__ uint32_t s_item;
    float4* sq = smem;
    float* ml_d = reinterpret_cast<float*>(smem + 8 * (size_t)Dq);  // [8 warps][8 queries][32]
    uint32_t* ml_r = reinterpret_cast<uint32_t*>(ml_d + 8 * 8 * 32);
    int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t n_items = counters[1];

    for (;;) {
        if (tid == 0) s_item = atomicAdd(work_counter, 1u);
        __syncthreads();
        uint32_t item = s_item;
        if (item >= n_items) break;
        ScanItem it = items[item];
        SegDesc sg = segs[it.seg];
        int nq = (int)it.nq;
        if (tid < nq) {
            uint2 e = seg_qlist[seg_qoff[it.seg] + it.qstart + tid];
            q_ids[tid] = e.x;
            q_slot[tid] = e.y;
        }
        __syncthreads();
        for (int idx = tid; idx < nq * Dq; idx += kSparseThreads) {
            int j = idx / Dq, c = idx - j * Dq;
            sq[idx] = xq4[(size_t)q_ids[j] * Dq + c];
        }
        __syncthreads();

        float my_d[8], thr[8];
        uint32_t my_r[8];
        float* alld[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            my_d[j] = __int_as_float(0x7f800000);
            thr[j] = __int_as_float(0x7f800000);
            my_r[j] = kNoRow;
            alld[j] = nullptr;
            if (ALLDIST && j < nq) alld[j] = alldist + (size_t)q_slot[j] * kSegVecs;
        }
        for (int g = warp; g < (int)sg.ng; g += 16) {
            int g2 = g + 8;
            bool has1 = g2 < (int)sg.ng;
            const float4* v0 = vecs + f4_index(sg.g0 + g, Dq, 0, 0);  // + c*kSuper + lane
            const float4* v1 = vecs + f4_index(sg.g0 + (has1 ? g2 : g), Dq, 0, 0);
            float a0[8], a1[8];
            switch (nq) {
                case 8: sparse_group_pair<8>(v0, v1, has1, Dq, sq, lane, a0, a1); break;
                case 7: sparse_group_pair<7>(v0, v1, has1, Dq, sq, lane, a0, a1); break;
                case 6: sparse_group_pair<6>(v0, v1, has1, Dq, sq, lane, a0, a1); break;
                case 5: sparse_group_pair<5>(v0, v1, has1, Dq, sq, lane, a0, a1); break;
                case 4: sparse_group_pair<4>(v0, v1, has1, Dq, sq, lane, a0, a1); break;
                case 3: sparse_group_pair<3>(v0, v1, has1, Dq, sq, lane, a0, a1); break;
                case 2: sparse_group_pair<2>(v0, v1, has1, Dq, sq, lane, a0, a1); break;
                default: sparse_group_pair<1>(v0, v1, has1, Dq, sq, lane, a0, a1); break;
            }
            sparse_select<ALLDIST>(a0, nq, lane, (sg.g0 + g) * 32u, (uint32_t)g * 32u, sg.nvalid, k, my_d, my_r, thr, alld);
            if (has1)
                sparse_select<ALLDIST>(a1, nq, lane, (sg.g0 + g2) * 32u, (uint32_t)g2 * 32u, sg.nvalid, k, my_d, my_r, thr,
                                       alld);
        }
        if (!ALLDIST) {
            // cross-warp merge: warp j merges query j's 8 partial lists by (dist, row)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (j < nq) {
                    ml_d[(warp * 8 + j) * 32 + lane] = my_d[j];
                    ml_r[(warp * 8 + j) * 32 + lane] = my_r[j];
                }
            }
            __syncthreads();
            if (warp < nq) {
                float fd = __int_as_float(0x7f800000);
                uint32_t fr = kNoRow;
                for (int w = 0; w < 8; w++) {
                    float d = ml_d[(w * 8 + warp) * 32 + lane];
                    uint32_t r = ml_r[(w * 8 + warp) * 32 + lane];
                    float td = __shfl_sync(kFull, fd, k - 1);
                    uint32_t tr = __shfl_sync(kFull, fr, k - 1);
                    unsigned m = __ballot_sync(kFull, r != kNoRow && lane < (int)k && (d < td || (d == td && r < tr)));
                    while (m) {
                        int src = __ffs(m) - 1;
                        m &= m - 1;
                        float cd = __shfl_sync(kFull, d, src);
                        uint32_t cr = __shfl_sync(kFull, r, src);
                        td = __shfl_sync(kFull, fd, k - 1);
                        tr = __shfl_sync(kFull, fr, k - 1);
                        if (cd < td || (cd == td && cr < tr)) warp_insert_lex(cd, cr, fd, fr, lane);
                    }
                }
                if (lane < (int)k) {
                    size_t o = (size_t)q_slot[warp] * k + lane;
                    cand_d[o] = fd;
                    cand_r[o] = fr;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < nq)
                    for (uint32_t pos = sg.ng * 32u + warp * 32u + lane; pos < (uint32_t)kSegVecs; pos += kSparseThreads)
                        alld[j][pos] = __int_as_float(0x7f800000);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------
// K5 second stage: per-query merge of its result slots (ordered by probe rank, then
// segment) into the final top-k; maps rows to external ids.  One warp per query.
// Equal distances keep slot order, i.e. probe-rank then list order (the order a stable
// sort over candidates gathered in probe order yields, src/ivf_index.rs:252-266).
// ------------------------------------------------------------------------------------
__global__ void merge_slots_kernel(const float* __restrict__ cand_d, const uint32_t* __restrict__ cand_r,
                                   const uint32_t* __restrict__ slot_off, uint32_t nq, uint32_t nprobe, uint32_t k,
                                   uint32_t kout, const uint64_t* __restrict__ row_ext, float* __restrict__ D,
                                   int64_t* __restrict__ I, uint32_t* __restrict__ out_rows) {
    uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (q >= nq) return;
    uint32_t s0 = slot_off[(size_t)q * nprobe], s1 = slot_off[(size_t)(q + 1) * nprobe];
    float fd = __int_as_float(0x7f800000);
    uint32_t fr = kNoRow;
    float thr = fd;
    for (uint32_t s = s0; s < s1; s++) {
        float d = __int_as_float(0x7f800000);
        uint32_t r = kNoRow;
        if (lane < (int)k) {
            d = cand_d[(size_t)s * k + lane];
            r = cand_r[(size_t)s * k + lane];
        }
        unsigned m = __ballot_sync(kFull, r != kNoRow && d < thr);
        while (m) {
            int src = __ffs(m) - 1;
            m &= m - 1;
            float cd = __shfl_sync(kFull, d, src);
            uint32_t cr = __shfl_sync(kFull, r, src);
            if (cd < thr) {
                warp_insert_stable(cd, cr, fd, fr, lane);
                thr = __shfl_sync(kFull, fd, k - 1);
            }
        }
    }
    if (lane < (int)k) {
        size_t o = (size_t)q * kout + lane;
        bool ok = fr != kNoRow;
        D[o] = ok ? fd : __int_as_float(0x7f800000);
        I[o] = ok ? (int64_t)row_ext[fr] : -1;
        if (out_rows) out_rows[o] = fr;
    }
}
// pad columns [k, kout) when the caller's k exceeds the clamped k
__global__ void pad_output_kernel(float* D, int64_t* I, uint32_t* rows, unsigned long long* keys, uint64_t nq, uint32_t k,
                                  uint32_t kout) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t w = kout - k;
    if (i >= nq * w) return;
    size_t q = i / w, t = k + i % w;
    D[q * kout + t] = __int_as_float(0x7f800000);
    I[q * kout + t] = -1;
    if (rows) rows[q * kout + t] = kNoRow;
    if (keys) keys[q * kout + t] = ~0ull;
}

// Large-k path: positions selected from the per-query all-distance rows -> ids.
// sel_pos[q*k + t] is a position in the query's concatenated slot rows.
__global__ void alldist_finish_kernel(const uint32_t* __restrict__ sel_pos, const float* __restrict__ sel_val,
                                      const uint32_t* __restrict__ slot_off, const uint32_t* __restrict__ slot_seg,
                                      const SegDesc* __restrict__ segs, uint32_t nprobe, uint64_t nq, uint32_t k,
                                      uint32_t kout, const uint64_t* __restrict__ row_ext, float* __restrict__ D,
                                      int64_t* __restrict__ I, uint32_t* __restrict__ out_rows,
                                      const uint32_t* __restrict__ slot_rank, const uint32_t* __restrict__ list_rowdelta,
                                      unsigned long long* __restrict__ out_keys) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * (size_t)k) return;
    size_t q = i / k;
    uint32_t t = (uint32_t)(i % k);
    uint32_t pos = sel_pos[i];
    size_t o = q * kout + t;
    if (pos == kNoRow) {
        D[o] = __int_as_float(0x7f800000);
        I[o] = -1;
        if (out_rows) out_rows[o] = kNoRow;
        if (out_keys) out_keys[o] = ~0ull;
        return;
    }
    uint32_t slot = slot_off[q * nprobe] + pos / kSegVecs;
    SegDesc sg = segs[slot_seg[slot]];
    uint32_t row = sg.g0 * 32u + pos % kSegVecs;
    D[o] = sel_val[i];
    I[o] = (int64_t)row_ext[row];
    if (out_rows) out_rows[o] = row;
    // (probe rank, global row): the order of the reference's stable sort over candidates gathered in probe order
    if (out_keys) out_keys[o] = ((unsigned long long)slot_rank[slot] << 32) | (uint32_t)(row + list_rowdelta[sg.list]);
}
__global__ void alldist_rows_kernel(const uint32_t* __restrict__ slot_off, uint32_t nprobe, uint64_t nq,
                                    uint64_t* __restrict__ row_off, uint32_t* __restrict__ row_len) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint32_t s0 = slot_off[q * nprobe], s1 = slot_off[(q + 1) * nprobe];
    row_off[q] = (uint64_t)s0 * kSegVecs;
    row_len[q] = (s1 - s0) * kSegVecs;
}

// Gather result payloads (include_vectors, src/api.rs:213-217) from the interleaved store.
__global__ void gather_vectors_kernel(const float* __restrict__ vecs, int Dq, int D, const uint32_t* __restrict__ rows,
                                      size_t nres, float* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nres * (size_t)D) return;
    size_t res = i / D;
    int d = (int)(i % D);
    uint32_t row = rows[res];
    float v = 0.0f;
    if (row != kNoRow) {
        v = vecs[(f4_row_base(row, Dq) + (size_t)(d >> 2) * kSuper) * 4 + (d & 3)];
    }
    out[i] = v;
}

// Interleave rows into groups: dst group layout from row-major src via a row map
// (row_src[r] = source vector index or kNoRow for padding).
__global__ void interleave_kernel(const float* __restrict__ src, int D, int Dq, const uint32_t* __restrict__ row_src,
                                  size_t nrows, float4* __restrict__ dst) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 of dst
    if (i >= nrows * (size_t)Dq) return;
    // dst index i = (supergroup * Dq + c) * 128 + r
    size_t sg = i / ((size_t)Dq * kSuper);
    size_t rem = i - sg * (size_t)Dq * kSuper;
    int c = (int)(rem >> 7), r = (int)(rem & 127);
    uint32_t s = row_src[sg * kSuper + r];
    float4 v = make_float4(0, 0, 0, 0);
    if (s != kNoRow) {
        const float* p = src + (size_t)s * D + 4 * c;
        int left = D - 4 * c;
        v.x = p[0];
        if (left > 1) v.y = p[1];
        if (left > 2) v.z = p[2];
        if (left > 3) v.w = p[3];
    }
    dst[i] = v;
}

// Multi-GPU: merge nruns per-rank results -- run r = D[nq][k] f32 at Dr + r * sD, I[nq][k] i64 at Ir + r * sI and,
// optionally, keys K[nq][k] u64 at Kr + r * sK (strides in elements) -- each sorted ascending by (distance, key) and
// padded with +inf, into the global top-k.  key = (probe rank << 32 | global row): (distance, key) is the order of the
// reference's stable sort over candidates gathered in probe order (ivf_index.rs:249-266), so the merged answer is
// bit-identical to the one-GPU answer however the index was split, ties included.  Without keys ties go to the lower
// run.  One thread per (query, run, position): its place in the merged order is its position in its own run plus, for
// every other run, the number of entries that sort before it (binary search).  Any k.
__global__ void merge_runs_kernel(const float* __restrict__ Dr, size_t sD, const int64_t* __restrict__ Ir, size_t sI,
                                  const unsigned long long* __restrict__ Kr, size_t sK, uint32_t nruns, uint64_t nq, uint32_t k,
                                  uint64_t per_group, float* __restrict__ D, int64_t* __restrict__ I) {
    // Query q belongs to query group g = q / per_group; its runs are runs g * nruns .. + nruns - 1, which hold the group's
    // queries only (row q - g * per_group).  One group (per_group >= nq) is the plain case.
    const size_t per = nq * (size_t)k;
    const float kInf = __int_as_float(0x7f800000);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per * nruns; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / per);
        const size_t e = i - (size_t)r * per, q = e / k;
        const uint32_t t = (uint32_t)(e - q * k);
        const size_t g = q / per_group, row = (q - g * per_group) * k, run0 = g * nruns;
        const float d = Dr[(run0 + r) * sD + row + t];
        if (!(d < kInf)) continue;  // padding: a returned distance is never +inf
        const unsigned long long key = Kr ? Kr[(run0 + r) * sK + row + t] : 0ull;
        uint32_t pos = t;
        for (uint32_t r2 = 0; r2 < nruns && pos < k; r2++) {
            if (r2 == r) continue;
            const float* d2 = Dr + (run0 + r2) * sD + row;
            const unsigned long long* k2 = Kr ? Kr + (run0 + r2) * sK + row : nullptr;
            uint32_t lo = 0, hi = k;  // first entry of run r2 that does NOT sort before (d, key)
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                const float dm = d2[mid];
                const bool before = dm < d || (dm == d && (k2 ? k2[mid] < key : r2 < r));
                if (before) lo = mid + 1;
                else hi = mid;
            }
            pos += lo;
        }
        if (pos < k) {
            D[q * k + pos] = d;
            I[q * k + pos] = Ir[(run0 + r) * sI + row + t];
        }
    }
}

// ====================================================================================
// host launchers
// ====================================================================================
static int num_sms() { return device_num_sms(); }

void launch_fill_u32(uint32_t* p, uint32_t v, size_t n, cudaStream_t st) {
    if (!n) return;
    fill_u32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(p, v, n);
    VIDX_LAUNCHED();
}
void launch_pad_rows(const float* in, float* out, uint64_t nrows, int D, int Dp, cudaStream_t st) {
    size_t n = nrows * (size_t)Dp;
    if (!n) return;
    pad_rows_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(in, out, nrows, D, Dp);
    VIDX_LAUNCHED();
}
void launch_interleave(const float* src, int D, int Dq, const uint32_t* row_src, size_t nrows, float4* dst,
                       cudaStream_t st) {
    size_t n = nrows * (size_t)Dq;
    if (!n) return;
    interleave_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(src, D, Dq, row_src, nrows, dst);
    VIDX_LAUNCHED();
}
void launch_coarse_dist(const float4* cents, int ngroups, int Dq, const float4* xq4, uint32_t nq, float* out, uint32_t ldo,
                        cudaStream_t st) {
    if (!nq || !ngroups) return;
    static PerDeviceSize attr;  // the opt-in is per device
    if (attr.needs(kDenseSmemBytes)) {
        VIDX_CUDA(cudaFuncSetAttribute(coarse_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmemBytes));
        attr.set(kDenseSmemBytes);
    }
    dim3 grid((unsigned)ceil_div(ngroups, kTileGroups), (unsigned)ceil_div(nq, kTileQ));
    coarse_dist_kernel<<<grid, kDenseThreads, kDenseSmemBytes, st>>>(cents, ngroups, Dq, xq4, nq, out, ldo);
    VIDX_LAUNCHED();
}
uint32_t select_kcap(uint32_t k) {
    uint32_t c = 32;
    while (c < k) c <<= 1;
    return c;
}
void launch_select_topk(const float* vals, const uint64_t* row_off, const uint32_t* row_len, uint64_t ld, uint32_t n_fixed,
                        uint64_t nrows, uint32_t k, uint32_t* out_pos, float* out_val, cudaStream_t st) {
    if (!nrows || !k) return;
    uint32_t kcap = select_kcap(k);
    size_t smem = (size_t)kcap * 8;
    if (smem > 200 * 1024) throw ApiError(6, "select_topk: k too large for this build (max 25600)");
    static PerDeviceSize attr_set;
    if (smem > 48 * 1024 && attr_set.needs(smem)) {
        VIDX_CUDA(cudaFuncSetAttribute(select_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set.set(smem);
    }
    select_topk_kernel<<<(unsigned)nrows, kSelThreads, smem, st>>>(vals, row_off, row_len, ld, n_fixed, k, kcap, out_pos,
                                                                    out_val);
    VIDX_LAUNCHED();
}
void launch_select_small(const float* vals, const uint64_t* row_off, const uint32_t* row_len, uint64_t ld, uint32_t n_fixed,
                         uint64_t nrows, uint32_t k, uint32_t* out_pos, float* out_val, cudaStream_t st) {
    if (!nrows || !k) return;
    if (k > 32) throw ApiError(6, "select_small: k > 32");
    select_small_kernel<<<(unsigned)ceil_div(nrows * 32, 256), 256, 0, st>>>(vals, row_off, row_len, ld, n_fixed, nrows, k, out_pos, out_val);
    VIDX_LAUNCHED();
}
void launch_group_count(const uint32_t* probes, size_t npairs, uint32_t nprobe, const uint2* list_seg,
                        const uint32_t* only_flag, uint32_t* pair_ns, uint32_t* seg_cnt, cudaStream_t st) {
    if (!npairs) return;
    group_count_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, st>>>(probes, npairs, nprobe, list_seg, only_flag, pair_ns,
                                                                         seg_cnt);
    VIDX_LAUNCHED();
}
void launch_group_fill(const uint32_t* probes, size_t npairs, uint32_t nprobe, const uint2* list_seg,
                       const uint32_t* slot_off, const uint32_t* seg_qoff, uint32_t* seg_cur, uint2* seg_qlist,
                       uint32_t* slot_seg, const uint32_t* only_flag, uint32_t* slot_rank, cudaStream_t st) {
    if (!npairs) return;
    group_fill_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, st>>>(probes, npairs, nprobe, list_seg, slot_off, seg_qoff,
                                                                        seg_cur, seg_qlist, slot_seg, only_flag, slot_rank);
    VIDX_LAUNCHED();
}
void launch_group_items(const uint32_t* seg_cnt, uint32_t nseg, uint32_t sparse_max, ScanItem* dense, ScanItem* sparse,
                        uint32_t* counters, cudaStream_t st) {
    if (!nseg) return;
    group_items_kernel<<<(unsigned)ceil_div(nseg, 256), 256, 0, st>>>(seg_cnt, nseg, sparse_max, dense, sparse, counters);
    VIDX_LAUNCHED();
}
size_t sparse_smem_bytes(int Dq) { return (size_t)8 * Dq * 16 + 8 * 8 * 32 * 8; }
bool sparse_supported(int Dq) { return sparse_smem_bytes(Dq) <= 100 * 1024; }

void launch_scan(bool alldist, const float4* vecs, int Dq, const float4* xq4, const SegDesc* segs,
                 const uint32_t* seg_qoff, const uint2* seg_qlist, const ScanItem* dense, const ScanItem* sparse,
                 const uint32_t* counters, uint32_t* work_counters, uint32_t k, float* cand_d, uint32_t* cand_r,
                 float* alld, bool use_sparse, cudaStream_t st) {
    static PerDeviceSize attr;
    if (attr.needs(1)) {
        VIDX_CUDA(cudaFuncSetAttribute(scan_dense_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmemBytes));
        VIDX_CUDA(cudaFuncSetAttribute(scan_dense_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmemBytes));
        VIDX_CUDA(cudaFuncSetAttribute(scan_sparse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        VIDX_CUDA(cudaFuncSetAttribute(scan_sparse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr.set(1);
    }
    int grid = num_sms() * 2;
    if (alldist)
        scan_dense_kernel<true><<<grid, kDenseThreads, kDenseSmemBytes, st>>>(vecs, Dq, xq4, segs, seg_qoff, seg_qlist, dense,
                                                                             counters, work_counters, k, cand_d, cand_r, alld);
    else
        scan_dense_kernel<false><<<grid, kDenseThreads, kDenseSmemBytes, st>>>(vecs, Dq, xq4, segs, seg_qoff, seg_qlist, dense,
                                                                              counters, work_counters, k, cand_d, cand_r, alld);
    VIDX_LAUNCHED();
    if (use_sparse) {
        size_t sm = sparse_smem_bytes(Dq);
        if (alldist)
            scan_sparse_kernel<true><<<grid, kSparseThreads, sm, st>>>(vecs, Dq, xq4, segs, seg_qoff, seg_qlist, sparse, counters,
                                                                      work_counters + 1, k, cand_d, cand_r, alld);
        else
            scan_sparse_kernel<false><<<grid, kSparseThreads, sm, st>>>(vecs, Dq, xq4, segs, seg_qoff, seg_qlist, sparse, counters,
                                                                       work_counters + 1, k, cand_d, cand_r, alld);
        VIDX_LAUNCHED();
    }
}
void launch_merge_slots(const float* cand_d, const uint32_t* cand_r, const uint32_t* slot_off, uint32_t nq, uint32_t nprobe,
                        uint32_t k, uint32_t kout, const uint64_t* row_ext, float* D, int64_t* I, uint32_t* out_rows,
                        cudaStream_t st) {
    if (!nq) return;
    merge_slots_kernel<<<(unsigned)ceil_div((size_t)nq * 32, 256), 256, 0, st>>>(cand_d, cand_r, slot_off, nq, nprobe, k, kout,
                                                                                  row_ext, D, I, out_rows);
    VIDX_LAUNCHED();
}
void launch_pad_output(float* D, int64_t* I, uint32_t* rows, unsigned long long* keys, uint64_t nq, uint32_t k, uint32_t kout,
                       cudaStream_t st) {
    if (kout <= k || !nq) return;
    size_t n = nq * (size_t)(kout - k);
    pad_output_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(D, I, rows, keys, nq, k, kout);
    VIDX_LAUNCHED();
}
void launch_alldist_rows(const uint32_t* slot_off, uint32_t nprobe, uint64_t nq, uint64_t* row_off, uint32_t* row_len,
                         cudaStream_t st) {
    if (!nq) return;
    alldist_rows_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(slot_off, nprobe, nq, row_off, row_len);
    VIDX_LAUNCHED();
}
void launch_alldist_finish(const uint32_t* sel_pos, const float* sel_val, const uint32_t* slot_off, const uint32_t* slot_seg,
                           const SegDesc* segs, uint32_t nprobe, uint64_t nq, uint32_t k, uint32_t kout,
                           const uint64_t* row_ext, float* D, int64_t* I, uint32_t* out_rows, const uint32_t* slot_rank,
                           const uint32_t* list_rowdelta, unsigned long long* out_keys, cudaStream_t st) {
    size_t n = nq * (size_t)k;
    if (!n) return;
    alldist_finish_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(sel_pos, sel_val, slot_off, slot_seg, segs, nprobe, nq, k,
                                                                      kout, row_ext, D, I, out_rows, slot_rank, list_rowdelta, out_keys);
    VIDX_LAUNCHED();
}
void launch_gather_vectors(const float* vecs, int Dq, int D, const uint32_t* rows, size_t nres, float* out, cudaStream_t st) {
    size_t n = nres * (size_t)D;
    if (!n) return;
    gather_vectors_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(vecs, Dq, D, rows, nres, out);
    VIDX_LAUNCHED();
}
void launch_merge_runs(const float* Dr, size_t sD, const int64_t* Ir, size_t sI, const unsigned long long* Kr, size_t sK,
                       uint32_t nruns, uint64_t nq, uint32_t k, float* D, int64_t* I, cudaStream_t st, uint64_t per_group) {
    if (!nq || !k) return;
    if (per_group == 0 || per_group > nq) per_group = nq;
    launch_pad_output(D, I, nullptr, nullptr, nq, 0, k, st);  // slots nobody claims: fewer than k candidates in all
    const size_t n = nq * (size_t)k * nruns;
    const unsigned grid = (unsigned)std::min<size_t>(ceil_div(n, 256), (size_t)num_sms() * 32);
    merge_runs_kernel<<<grid, 256, 0, st>>>(Dr, sD, Ir, sI, Kr, sK, nruns, nq, k, per_group, D, I);
    VIDX_LAUNCHED();
}

}  // namespace vidx

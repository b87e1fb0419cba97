// common.cuh -- shared host/device helpers for libvidx_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace vidx {

constexpr int kGroup = 32;          // vectors per group (one per lane)
constexpr int kSuper = 128;         // vectors per supergroup = 4 groups: the unit of the HBM layout and of a tensor-core tile
constexpr int kSegGroups = 32;      // groups per segment (<= 1024 vectors): the scan work unit
constexpr int kSegVecs = kGroup * kSegGroups;
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kNoRow = 0xffffffffu;

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};
struct ApiError : std::runtime_error {
    int code;
    ApiError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define VIDX_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            char _b[512];                                                                        \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                     __LINE__, cudaGetErrorString(_e));                                          \
            throw ::vidx::CudaError(_b);                                                         \
        }                                                                                        \
    } while (0)

extern std::atomic<uint64_t> g_kernel_launches;
#define VIDX_LAUNCHED()                            \
    do {                                           \
        ::vidx::g_kernel_launches.fetch_add(1);    \
        VIDX_CUDA(cudaGetLastError());             \
    } while (0)

// Allocations made by this thread's DevBufs so far: a cached CUDA graph of a search holds the addresses of its context's
// buffers, so a lease that saw one of them move drops the graph (index.cu, CtxLease).
inline uint64_t& devbuf_allocs() {
    static thread_local uint64_t n = 0;
    return n;
}

// Grow-only device buffer.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        release();
        size_t want = bytes + bytes / 8 + 256;
        devbuf_allocs()++;
        VIDX_CUDA(cudaMalloc(&p, want));
        cap = want;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        VIDX_CUDA(cudaMallocHost(&p, bytes + 256));
        cap = bytes + 256;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// ---- per-device one-time state ---------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to a DEVICE, not to the process: a
// second handle on another GPU of the same process needs its own opt-ins.  Slots are indexed by the current
// device ordinal; racing threads at worst set the same attribute twice.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    VIDX_CUDA(cudaGetDevice(&dev));
    return dev < 0 || dev >= kMaxDevices ? 0 : dev;
}
struct PerDeviceSize {
    std::atomic<size_t> v[kMaxDevices];
    PerDeviceSize() { for (auto& x : v) x.store(0); }
    // true when `want` exceeds what this device has been configured for so far (the caller then sets the attribute)
    bool needs(size_t want) const { return want > v[current_device()].load(std::memory_order_acquire); }
    void set(size_t want) { v[current_device()].store(want, std::memory_order_release); }
};
inline int device_num_sms() {
    static std::atomic<int> sms[kMaxDevices];
    const int dev = current_device();
    int n = sms[dev].load(std::memory_order_acquire);
    if (!n) {
        VIDX_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        sms[dev].store(n, std::memory_order_release);
    }
    return n;
}

// One segment = up to kSegGroups consecutive groups of one list.
struct SegDesc {
    uint32_t g0;      // first group (global group index into the interleaved store)
    uint32_t ng;      // groups in this segment
    uint32_t nvalid;  // valid vectors in this segment (<= ng*32)
    uint32_t list;    // owning list
};

// ---- HBM layout of vectors ------------------------------------------------------------
// A supergroup holds 128 consecutive vectors of one list as float4 [Dq][128]: the 16-byte
// chunk c (dims 4c..4c+3) of all 128 vectors is one contiguous 2 KB block, and any run of
// chunks of a supergroup is contiguous too.  That is at once
//   * coalesced for a warp (lane <-> vector: 512 contiguous bytes per chunk per group), and
//   * the tcgen05 no-swizzle K-major operand layout (8-row x 16-byte core matrices, SBO 128 B,
//     LBO 2048 B), so a K-slice of a list tile is ONE contiguous bulk copy into shared memory.
// Rows are numbered linearly (row = group*32 + lane); lists start on supergroup boundaries.
__host__ __device__ inline size_t f4_index(size_t group, int Dq, int c, int lane) {
    return ((group >> 2) * (size_t)Dq + c) * kSuper + (group & 3) * kGroup + lane;
}
__host__ __device__ inline size_t f4_row_base(size_t row, int Dq) {  // + c*kSuper per chunk
    return (row >> 7) * (size_t)Dq * kSuper + (row & 127);
}

#ifdef __CUDACC__
// ---- reference arithmetic ---------------------------------------------------------
// utils.rs:28-30: acc + (x - y)*(x - y) with separate roundings (Rust never fuses).
__device__ __forceinline__ float sqdiff_acc(float acc, float x, float y) {
    float d = __fsub_rn(x, y);
    return __fadd_rn(acc, __fmul_rn(d, d));
}
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- warp-resident sorted list: lane t holds the t-th best (dist, row) -------------
// STABLE insertion: the candidate goes after every entry with dist <= cd (arrival order
// breaks ties, as the reference's stable sort over candidates in scan order does,
// ivf_index.rs:265).
__device__ __forceinline__ void warp_insert_stable(float cd, uint32_t cr, float& my_d, uint32_t& my_r, int lane) {
    unsigned m = __ballot_sync(kFull, my_d <= cd);
    int pos = __popc(m);
    float up_d = __shfl_up_sync(kFull, my_d, 1);
    uint32_t up_r = __shfl_up_sync(kFull, my_r, 1);
    if (lane == pos) { my_d = cd; my_r = cr; }
    else if (lane > pos) { my_d = up_d; my_r = up_r; }
}
// LEX insertion: ordered by (dist, row); used where candidates do not arrive in row order.
__device__ __forceinline__ void warp_insert_lex(float cd, uint32_t cr, float& my_d, uint32_t& my_r, int lane) {
    unsigned m = __ballot_sync(kFull, (my_d < cd) || (my_d == cd && my_r < cr));
    int pos = __popc(m);
    float up_d = __shfl_up_sync(kFull, my_d, 1);
    uint32_t up_r = __shfl_up_sync(kFull, my_r, 1);
    if (lane == pos) { my_d = cd; my_r = cr; }
    else if (lane > pos) { my_d = up_d; my_r = up_r; }
}
#endif

}  // namespace vidx

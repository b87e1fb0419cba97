// Microbenchmark: issue cadence of tcgen05.mma kind::f16 (M=128, K=16) from ONE thread, for N = 128 / 256, with the A
// operand in shared memory (SS) or in tensor memory (TS), optionally with a concurrent stream of bulk copies into shared
// memory (what the scan's producer does).  Prints cycles per MMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
    asm volatile("{.reg .b64 da, db; .reg .pred p; setp.ne.b32 p, 1, 0; mov.b64 da, {%1, %3}; mov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_t, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
    asm volatile("{.reg .b64 db; .reg .pred p; setp.ne.b32 p, 1, 0; mov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;}" ::"r"(d), "r"(a_t), "r"(b_lo), "r"(hi), "r"(idesc) : "memory");
}
__global__ void __launch_bounds__(128, 1) k(int iters, int N, int ts, int copies, const unsigned char* src, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tb;
    __shared__ __align__(8) uint64_t bar, cbar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // halfs 1.0
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi = (128u >> 4) | (1u << 14), lbo = (2048u >> 4) << 16;
    const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3fffu) | lbo;                 // A: 32 KB at 0
    const uint32_t b_lo = ((smem_u32(smem + 32768) >> 4) & 0x3fffu) | (N == 256 ? ((4096u >> 4) << 16) : lbo);  // B: at 32 KB
    if (warp == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            if (lane == 0) {
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {
                    if (ts) mma_ts(tb + (N == 256 ? 0 : (i & 1) * 256), tb + 384 + 0 * ks, b_lo + ks * (N == 256 ? 512 : 256), hi, idesc);
                    else mma_ss(tb + (i & 1) * 256, a_lo + ks * 256, b_lo + ks * (N == 256 ? 512 : 256), hi, idesc);
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
            cyc[blockIdx.x] = clock64() - t0;
        }
    } else if (warp == 1 && copies) {
        // a stream of 32 KB bulk copies into the upper part of shared memory while the MMAs run
        if (lane == 0) {
            for (int i = 0; i < copies; i++) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar)), "r"(32768) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(smem + 98304 + (i & 1) * 32768)),
                             "l"(src + (size_t)((blockIdx.x * 64 + (i & 63)) * 32768)), "r"(32768), "r"(smem_u32(&cbar))
                             : "memory");
                uint32_t done = 0;
                while (!done)
                    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&cbar)), "r"(i & 1) : "memory");
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}
int main() {
    long long* cyc; unsigned char* src;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&src, (size_t)148 * 64 * 32768); cudaMemset(src, 0, (size_t)148 * 64 * 32768);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 168 * 1024);
    for (int N : {128, 256}) for (int ts : {0, 1}) for (int copies : {0, 1}) {
        const int iters = 4000;
        const int ncp = copies ? (N == 128 ? iters : 2 * iters) : 0;  // one 32 KB copy per 128 vectors of B, like the scan
        k<<<148, 128, 168 * 1024>>>(10, N, ts, copies ? 10 : 0, src, cyc); cudaDeviceSynchronize();
        k<<<148, 128, 168 * 1024>>>(iters, N, ts, ncp, src, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
        printf("N=%d A in %s, bulk copies %s: %.1f cycles per MMA (%.0f per 128 vectors x 128 dims) (%s)\n", N, ts ? "TMEM" : "smem", copies ? "on " : "off",
               (double)h[0] / (iters * 8.0), (double)h[0] / iters / (N / 128), cudaGetErrorString(e));
    }
    return 0;
}

// Microbenchmark: TMEM read bandwidth of tcgen05.ld (32x32b.x32) per SM, for 4 and 8 reading warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float ld32(uint32_t taddr) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) x ^= r[i];
    return __uint_as_float(x);
}
__global__ void k(int iters, int inflight, float* out, long long* cyc) {
    __shared__ uint32_t tb;
    int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tb)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    float acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        acc += ld32(base + ((i * 32) & 63) + ((i >> 1) & 3) * 128);
        if (inflight) acc += ld32(base + (((i + 1) * 32) & 63) + ((i >> 1) & 3) * 128);
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {1, 4, 8}) for (int grid : {1, 148}) for (int two : {0, 1}) {
        int iters = 20000;
        k<<<grid, warps * 32>>>(100, two, out, cyc); cudaDeviceSynchronize();
        k<<<grid, warps * 32>>>(iters, two, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double bytes = (double)iters * (two ? 2 : 1) * warps * 4096.0;
        printf("warps %d grid %3d ld/iter %d: %lld cycles, %.1f B/clk/SM, %.1f clk per 4KB ld per warp (%s)\n", warps, grid, two + 1, h[0], bytes / h[0],
               (double)h[0] / (iters * (two ? 2 : 1)), cudaGetErrorString(e));
    }
    return 0;
}

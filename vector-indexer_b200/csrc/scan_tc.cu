// scan_tc.cu -- tensor-core list scan for sm_100a: tcgen05 FP16 pre-filter + exact re-check.
//
// The inverted-list scan (src/ivf_index.rs:252-266) is, for a batch of queries grouped by
// probed list, the contraction  dot[q][v] = sum_d q[d]*v[d]  followed by a top-k.  The
// reference's distance is fp32 and results must be bit-exact, so the tensor cores are used
// as a FILTER only:
//
//   L(q,v) = (1-eps)*(|q|^2 + |v|^2) - c_abs - 2*dot_fp16(q,v)    is a guaranteed LOWER bound of
//   the reference distance: operands are rounded to fp16 (11 significant bits, |error| <= 2^-11
//   relative after a power-of-two scaling that keeps every component far inside fp16 range), so
//   |2 dot - 2 dot_fp16| <= 2^-10 |q||v| <= 2^-11 (|q|^2+|v|^2) plus c_abs for components that fall into the
//   fp16 subnormal range; eps = 1.5e-3 also covers every fp32 rounding involved.
//
//   A candidate survives iff L <= U, where U is an UPPER bound of the query's exact k-th best
//   distance: the k-th smallest L seen so far plus 2*eps*(|q|^2 + max|v|^2) + 2*c_abs, or a bound
//   published by another CTA for the same query.  Every true top-k member survives.
//
//   Survivors (a few dozen per query) get the reference's exact sequential fp32 distance in
//   finalize_kernel, which then selects the top-k by (distance, probe rank, row) -- the same
//   keys the exact scan kernels order by.  The final answer is bit-identical to the exact path.
//
// The index keeps an fp16 SHADOW of the vectors for this filter (same supergroup layout, 8 dims per 16-byte
// chunk); the fp32 store stays the source of every distance that is returned.
//
// Kernel anatomy (one CTA per SM, persistent, 12 warps -- ptxas sizes registers for whole warpgroups, so 384 threads get 168
// registers each and the epilogue's per-tile loop neither spills nor rematerialises; a work item = 128 queries x up to 512
// list tiles):
//   warp 0 producer  : one cp.async.bulk (1-D TMA) per K-slice of a 128-vector tile -- the HBM layout keeps
//                      a tile's chunks contiguous, so a list chunk is ONE linear stream -- into a ring of 32 KB
//                      stages, plus the tile's norm terms; completion on mbarriers (complete_tx::bytes)
//   warps 9, 10 MMA  : one elected thread issues tcgen05.mma.cta_group::1.kind::f16, M=128 queries x N=128
//                      vectors x K=16 per instruction; operands are read from shared memory through no-swizzle
//                      K-major descriptors: the chunk-major layout IS the UMMA core-matrix layout (8 rows x 16 B
//                      contiguous, SBO = 128 B, LBO = 2048 B).  A last K-step multiplies (a,a,a,0..) by the
//                      three fp16 terms of the row norm, so the accumulator holds the filter value itself.
//                      Two independent pipelines (issuer + half of the ring + two accumulator stages) alternate
//                      tiles: one thread issues an N = 128 MMA only every ~125 cycles (tools/micro/mma_rate.cu).
//   warps 1-8 epilogue: tcgen05.ld the 128x128 fp32 accumulator tile from TMEM (4 stages, 512 columns); two groups of
//                      four warps (one per TMEM lane quarter) alternate tiles, each thread owns one query row: one 3-input
//                      min per three columns, one branch per 32; hits go to a shared-memory queue
//   warp 11 selector : the rows' top-k sets and bounds, the hit queue, staged appends of survivors to the per-query lists
// Accumulators never leave the SM; HBM sees each list tile once per 128-query tile.
//
// Two launches of the one kernel: mode 2 = bounds pass (the epilogue only records the minimum of every 32 columns; the
// k smallest minima of a query are its first top-k set), mode 0 = main pass (hit path, queues, selector; "frozen" after
// a bounds pass over everything the query probes: bounds are final, survivors are only collected).
#include <cuda_fp16.h>

#include <algorithm>

#include "scan_tc.h"
#include "search.h"

namespace vidx {

constexpr int kTcThreads = 384;       // warp 0 producer; 1-8 epilogue; 9, 10 MMA (even / odd tiles); 11 selector.
                                      // 12 warps, not 16: ptxas sizes registers for whole warpgroups, so 384 threads get 168 registers
                                      // per thread instead of 128 for the epilogue's per-tile loop
constexpr int kTcEpiWarps = 8;
constexpr int kTcSelectors = 1;      // selector warps: each owns 32 query rows (one TMEM lane quarter), with its own hit queue and survivor staging
constexpr int kTcM = 128;            // queries per tile (UMMA M)
constexpr int kTcSeedRows = 128;     // seeding pass: queries per work item (every row starts cold there; the four selectors share the flood)
#ifndef VIDX_CHUNK_TILES
#define VIDX_CHUNK_TILES 512
#endif
#ifndef VIDX_ITEMS_PER_SM
#define VIDX_ITEMS_PER_SM 4
#endif
#ifndef VIDX_MIN_CHUNK_TILES
#define VIDX_MIN_CHUNK_TILES 8
#endif
constexpr int kTcMaxChunkTiles = VIDX_CHUNK_TILES; // tiles per work item: chosen on the device, 8..512 (1024..65536 vectors); an item
                                      // costs ~8-16 us beyond its tiles (set-up, pipeline fill and drain, merge of the rows' sets):
                                      // measured 2.30 / 2.17 / 2.09 ms per batch with at most 128 / 256 / 512 tiles per item
// hit queue entries (power of two) and survivors staged by the selector before a bulk append; the 32-entry
// top-k sets (k > 16) leave room for smaller ones only
__host__ __device__ constexpr int tc_queue_cap(int kr) { return kr > 16 ? 128 : 512; }
__host__ __device__ constexpr int tc_stage_cap(int kr) { return kr > 16 ? 144 : 512; }
constexpr int kTcAccStages = 4;      // accumulator tiles in TMEM
#ifndef VIDX_FLOOD_AFTER
#define VIDX_FLOOD_AFTER 2
#endif
constexpr uint32_t kTcFloodAfter = VIDX_FLOOD_AFTER;  // values a thread queues per item before it starts tracking its own k smallest
constexpr int kTcMergeTries = 4;     // end-of-item merge of a row's set into the shared one: attempts at the row's lock
constexpr uint32_t kTcACol = 384;    // A-in-TMEM variant: three accumulator stages, then the query tile (4 columns per 8 dims), then
                                     // the 8 columns of the norm step's A operand
constexpr int kTcTmemCols = 512;     // 4 accumulator stages x 128 columns
constexpr int kTcStages = 4;         // shared-memory ring: stages of 128 vectors x 128 dims of fp16 (32 KB), two per tile pipeline
constexpr int kTcStageChunks = 16;   // 16-byte chunks (8 halfs) of every vector per stage
constexpr int kTcStreamChunks = 8;   // ... when the query tile is streamed through the ring too (large D): a stage then holds
                                     // 16 KB of list chunks, the norm chunk and 16 KB of query chunks -- the same 34 KB
constexpr uint32_t kTcStageData = kTcStageChunks * kTcTileGroups * 512;  // 32 KB of vector chunks
constexpr uint32_t kTcStageBytes = kTcStageData + 2048;                  // + the tile's norm chunk (used by the last K-slice)
constexpr float kTcEps = 1.5e-3f;    // see header comment; needed: ~1.03e-3

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait suspends the thread for a hardware time slice per attempt; a wait that is still
// pending after ~2^24 attempts is a protocol bug, so trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// One lane of a converged warp (CUTLASS's elect_one_sync): lets the warp run a role's loop with warp-uniform
// values -- which the compiler keeps in uniform registers -- while a single thread issues.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}"
        : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Tiled TMA copy of one box of a 2-D tensor map into shared memory (coordinates innermost first).
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
// ---- CTA pair (cluster of two CTAs on one TPC, tcgen05 cta_group::2) ------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, uint4 v) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// the commit of a cta_group::2 MMA arrives on the barrier at the same offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// M = 256 (128 rows from each CTA of the pair), N = 128 (64 rows of B from each CTA), issued by one thread of the leader
template <bool ACC>
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, FP16 inputs, FP32 accumulate.  The two descriptors are given as 32-bit
// halves: only the 14-bit start-address field of the low word differs between the MMAs of a tile, so the
// issuing thread does one add per descriptor.
template <bool ACC>
__device__ __forceinline__ void tc_mma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
// The same with the A operand in tensor memory (row i in lane i, the 16 halfs of a K step in 8 consecutive columns): the query
// tile is then read from TMEM, not from shared memory -- whose bandwidth (128 B/clk) the B operand (4 KB per MMA) and the bulk
// copies already take most of.
template <bool ACC>
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(desc_hi), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
}
// 32 lanes x 8 columns from registers: thread t of the warp writes lane (quarter*32 + t)
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (quarter*32 + t), 32 consecutive columns.
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
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
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
// The same in two halves: issue only, and a wait that NAMES the destination registers -- the compiler then cannot move
// a use of them above the wait (a plain `asm volatile("tcgen05.wait::ld")` orders nothing with respect to registers).
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
}
__device__ __forceinline__ void tc_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
// Two 32-column loads in flight, one wait.
__device__ __forceinline__ void tc_ld32x2(uint32_t taddr0, uint32_t taddr1, float (&v)[64]) {
    uint32_t r[64];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%65];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr0), "r"(taddr1)
        : "memory");
#pragma unroll
    for (int i = 0; i < 64; i++) v[i] = __uint_as_float(r[i]);
}
// Shared-memory matrix descriptor, no swizzle, K-major: 8-row x 16-byte core matrices;
// LBO = byte distance between the two 16-byte K chunks of one instruction, SBO = byte distance
// between consecutive 8-row groups; version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// Instruction descriptor: D = F32 (bits 4-5 = 1), A = B = F16 (0 at bits 7-9 / 10-12), both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// norms, scales and the fp16 shadow store
// ------------------------------------------------------------------------------------------
int tc_dh(int D) { return 2 * ((D + 15) / 16); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// |v|^2 per stored row and the largest |component| of the store.
__global__ void row_norm_kernel(const float4* __restrict__ vecs, int Dq, const uint32_t* __restrict__ row_src, size_t nrows,
                                float* __restrict__ vn_true, uint32_t* __restrict__ stats) {
    size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.0f, m = 0.0f;
    if (row < nrows && row_src[row] != kNoRow) {
        const float4* p = vecs + f4_row_base(row, Dq);
        for (int c = 0; c < Dq; c++) {
            float4 v = p[(size_t)c * kSuper];
            s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
    }
    if (row < nrows) vn_true[row] = s;
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(&stats[0], __float_as_uint(m));
}
// Shadow store: chunk c of a row = dims 8c..8c+7 as fp16 (round to nearest) of v * 2^sv, zero beyond the dimension.
// Norm terms: b = (1-eps)|v|^2 * 2^(2sv-g) as three fp16 values rounded DOWN (hi + mid + lo <= b, so the filter
// value can only get smaller); NaN for padding rows (NaN never passes a '<=' test).
// perm != nullptr: shadow row `row` holds the vector of store row perm[row] (rows sorted by norm inside every 1024-vector segment,
// see Index::finish_store); vnorm32[row] = the same norm term as ONE fp32 value rounded down (epilogue-side norms, scan_tc_kernel<.., NB>).
__global__ void convert16_kernel(const float4* __restrict__ vecs, int Dq, int Dh, const uint32_t* __restrict__ row_src, size_t nrows,
                                 const float* __restrict__ vn_true, float vscale, float nscale, const uint32_t* __restrict__ perm,
                                 uint4* __restrict__ vecs16, uint4* __restrict__ vnorm, float* __restrict__ vnorm32) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // (row, chunk), row fastest: coalesced on both sides
    size_t row = i % nrows;
    int c = (int)(i / nrows);
    if (c >= Dh) return;
    const size_t srow = perm ? perm[row] : row;  // row of the fp32 store
    const bool pad = row_src[srow] == kNoRow;
    const float4* p = vecs + f4_row_base(srow, Dq);
    float4 a = make_float4(0, 0, 0, 0), b = a;
    if (!pad && 2 * c < Dq) a = p[(size_t)(2 * c) * kSuper];
    if (!pad && 2 * c + 1 < Dq) b = p[(size_t)(2 * c + 1) * kSuper];
    uint4 o;
    o.x = pack_h2(a.x * vscale, a.y * vscale);
    o.y = pack_h2(a.z * vscale, a.w * vscale);
    o.z = pack_h2(b.x * vscale, b.y * vscale);
    o.w = pack_h2(b.z * vscale, b.w * vscale);
    vecs16[((row >> 7) * (size_t)Dh + c) * kSuper + (row & 127)] = o;
    if (c == 0) {
        uint4 n = make_uint4(0, 0, 0, 0);
        if (pad) {
            n.x = 0x7e00u;  // fp16 NaN in the hi term
        } else {
            const float bv = (1.0f - kTcEps) * vn_true[srow] * nscale;
            const __half hi = __float2half_rd(bv);
            const float r1 = bv - __half2float(hi);
            const __half mid = __float2half_rd(r1);
            const float r2 = r1 - __half2float(mid);
            const __half lo = __float2half_rd(r2);
            n.x = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(mid) << 16);
            n.y = (uint32_t)__half_as_ushort(lo);
        }
        vnorm[row] = n;
        // (NaN for padding rows: their accumulator is 0, and NaN never passes a '<=' test -- not even against a cold row's +inf bound)
        if (vnorm32) vnorm32[row] = pad ? __int_as_float(0x7fc00000) : __fmul_rd(__fmul_rd(1.0f - kTcEps, vn_true[srow]), nscale);
    }
}
// The smallest norm term of every group of 32 shadow rows (fminf drops the NaNs of padding rows; NaN for a group of padding only): scan_tc_kernel<.., NB> compares a
// row's minimum over the group's 32 accumulator columns with its bound minus this.
__global__ void group_min_norm_kernel(const float* __restrict__ vnorm32, size_t ngroups, float* __restrict__ gmin) {
    const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= ngroups) return;
    float m = vnorm32[g * 32 + (threadIdx.x & 31)];
    for (int o = 16; o; o >>= 1) m = fminf(m, __shfl_xor_sync(kFull, m, o));
    if ((threadIdx.x & 31) == 0) gmin[g] = m;
}
__global__ void query_norm_kernel(const float4* __restrict__ xq4, int Dq, uint32_t nq, uint32_t k, float* __restrict__ qn,
                                  uint32_t* __restrict__ gthr_bits, uint32_t* __restrict__ cand_cnt, uint32_t* __restrict__ overflow,
                                  float* __restrict__ gtop, uint32_t* __restrict__ glock, uint32_t* __restrict__ stats) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.0f, m = 0.0f;
    if (q < nq) {
        for (int c = 0; c < Dq; c++) {
            float4 v = xq4[(size_t)q * Dq + c];
            s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
        qn[q] = s;
        gthr_bits[q] = 0x7f800000u;  // +inf
        cand_cnt[q] = 0;
        // a query with a NaN / infinite component has no usable filter values (fmaxf below drops NaN, so the batch scale
        // would not notice): the exact kernels answer it, as they do in scan mode 1
        overflow[q] = (s != s || isinf(s)) ? 1u : 0u;
        glock[q] = 0;
        glock[nq + q] = 0;  // seqlock version of gtop
        for (uint32_t i = 0; i < k; i++) gtop[(size_t)q * k + i] = __int_as_float(0x7f800000);
    }
    float sm = s;
    for (int o = 16; o; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
        sm = fmaxf(sm, __shfl_xor_sync(kFull, sm, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (m > 0.0f) atomicMax(&stats[0], __float_as_uint(m));
        if (sm > 0.0f) atomicMax(&stats[1], __float_as_uint(sm));
    }
}
// The batch's query scale: largest |2q component| * 2^sq in [2^6, 2^7), moved if the norm partner a = 2^(sq-sv+g)
// would leave fp16's normal range.
__global__ void tc_scale_kernel(const uint32_t* __restrict__ qstats, int sv, int g, int D, float vmax, float vn_max,
                                TcScale* __restrict__ out) {
    const float qmax2 = 2.0f * __uint_as_float(qstats[0]);
    const float qn_max = __uint_as_float(qstats[1]);
    int e = 0;
    if (qmax2 > 0.0f) frexpf(qmax2, &e);  // qmax2 = m * 2^e, m in [0.5, 1)
    int sq = qmax2 > 0.0f ? 7 - e : 0;
    bool ok = isfinite(qmax2) && isfinite(qn_max);
    int ea = sq - sv + g;
    if (ea > 15) { sq -= ea - 15; ea = 15; }            // smaller query scale: always safe (precision only)
    if (ea < -14) {
        const int up = -14 - ea;                          // larger query scale: must stay below fp16 overflow
        if (up <= 8) { sq += up; ea = -14; } else ok = false;
    }
    if (sq > 100 || sq < -100) ok = false;
    TcScale t;
    t.S = ok ? ldexpf(1.0f, sq + sv) : 1.0f;
    t.invS = ok ? ldexpf(1.0f, -(sq + sv)) : 1.0f;
    t.qmul = ok ? -ldexpf(1.0f, sq + 1) : 0.0f;
    t.a_ones = ok ? ldexpf(1.0f, ea) : 0.0f;
    // components below fp16's normal range carry an absolute error of at most 2^-25 (scaled); see the header
    t.c_abs = ok ? sqrtf((float)D) * (ldexpf(1.0f, -25 - sq) * sqrtf(vn_max) * 2.0f + ldexpf(1.0f, -24 - sv) * sqrtf(qn_max)) : 0.0f;
    t.ok = ok ? 1u : 0u;
    *out = t;
    (void)vmax;
}

// ------------------------------------------------------------------------------------------
// grouping by list
// ------------------------------------------------------------------------------------------
// max_rank != 0: only each query's max_rank nearest lists (the seeding pass).
// (a few giant lists take most pairs: the lanes of a warp that hit the same list share one atomic)
__global__ void tc_count_kernel(const uint32_t* __restrict__ probes, size_t npairs, uint32_t nprobe, uint32_t max_rank,
                                const uint2* __restrict__ list_seg, uint32_t* __restrict__ list_cnt) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool ok = p < npairs && !(max_rank && (p % nprobe) >= max_rank);
    uint32_t l = ok ? probes[p] : kNoRow;
    ok = ok && l != kNoRow;
    if (ok) {
        uint2 sr = list_seg[l];
        ok = sr.y > sr.x;  // owned, non-empty list
    }
    const unsigned okm = __ballot_sync(kFull, ok);
    if (!ok) return;
    const unsigned m = __match_any_sync(okm, l);
    if (lane == __ffs(m) - 1) atomicAdd(&list_cnt[l], (uint32_t)__popc(m));
}
__global__ void tc_fill_kernel(const uint32_t* __restrict__ probes, size_t npairs, uint32_t nprobe, uint32_t max_rank,
                               const uint2* __restrict__ list_seg, const uint32_t* __restrict__ list_qoff,
                               uint32_t* __restrict__ list_cur, uint2* __restrict__ list_qlist) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool ok = p < npairs && !(max_rank && (p % nprobe) >= max_rank);
    uint32_t l = ok ? probes[p] : kNoRow;
    ok = ok && l != kNoRow;
    if (ok) {
        uint2 sr = list_seg[l];
        ok = sr.y > sr.x;
    }
    const unsigned okm = __ballot_sync(kFull, ok);
    if (!ok) return;
    const unsigned m = __match_any_sync(okm, l);
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&list_cur[l], (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    const uint32_t i = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
    list_qlist[list_qoff[l] + i] = make_uint2((uint32_t)(p / nprobe), (uint32_t)(p % nprobe));
}
// Work items of list l = (#query tiles) x (#vector chunks), enumerated chunk-major so that CTAs
// running at the same time share a vector chunk (L2 reuse) and a query's later chunks start with a
// warm bound.  The chunk length is chosen on the device from the total tile count so that a batch
// yields several items per SM: ctl[0] = total (query tile x vector tile) count, ctl[1] = chunk tiles.
__global__ void tc_work_kernel(const uint32_t* __restrict__ list_cnt, const uint32_t* __restrict__ list_ngroups, uint32_t nlist,
                               uint32_t qrows, unsigned long long* __restrict__ total) {
    // (main pass only; the seeding pass has one fixed-size chunk per list)
    uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlist) return;
    uint32_t c = list_cnt[l];
    if (!c) return;
    uint32_t ntiles = (list_ngroups[l] + kTcTileGroups - 1) / kTcTileGroups;
    atomicAdd(total, (unsigned long long)((c + qrows - 1) / qrows) * ntiles);
}
__global__ void tc_items_kernel(const uint32_t* __restrict__ list_cnt, const uint32_t* __restrict__ list_ngroups, uint32_t nlist,
                                const unsigned long long* __restrict__ total, uint32_t num_sms, uint32_t seed_tiles,
                                uint32_t items_per_sm, uint32_t min_chunk, uint32_t qrows_main, uint32_t* __restrict__ chunk_out,
                                uint32_t* __restrict__ items_per_list) {
    uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t chunk;
    if (seed_tiles) {
        chunk = min(seed_tiles, 16u);  // the bounds launch's items are independent: short chunks spread them over the SMs
    } else {
        unsigned long long per = *total / ((unsigned long long)items_per_sm * num_sms);
        chunk = (uint32_t)min((unsigned long long)kTcMaxChunkTiles, max((unsigned long long)min_chunk, per));
    }
    if (l == 0) *chunk_out = chunk;
    if (l >= nlist) return;
    uint32_t c = list_cnt[l];
    uint32_t ntiles = (list_ngroups[l] + kTcTileGroups - 1) / kTcTileGroups;
    if (seed_tiles) ntiles = min(ntiles, seed_tiles);
    uint32_t nch = (ntiles + chunk - 1) / chunk;
    const uint32_t qrows = seed_tiles ? (uint32_t)kTcSeedRows : qrows_main;
    items_per_list[l] = c ? ((c + qrows - 1) / qrows) * nch : 0u;
}

// One record per work item, so that a CTA decodes an item with a single 32-byte load (prefetched one item ahead)
// instead of a binary search and a chain of dependent loads.
__global__ void tc_expand_kernel(const uint32_t* __restrict__ list_cnt, const uint32_t* __restrict__ list_ngroups,
                                 const uint32_t* __restrict__ list_g0, const uint32_t* __restrict__ list_qoff,
                                 const uint32_t* __restrict__ item_off, const uint32_t* __restrict__ chunk_tiles, uint32_t nlist,
                                 uint32_t seed_tiles, uint32_t qrows_main, const uint32_t* __restrict__ a_rowoff,
                                 TcItem* __restrict__ items) {
    const uint32_t l = blockIdx.x;
    if (l >= nlist) return;
    const uint32_t i0 = item_off[l], n = item_off[l + 1] - i0;
    if (!n) return;
    const uint32_t cnt = list_cnt[l], ngl = list_ngroups[l], g_list = list_g0[l], qoff = list_qoff[l], chunk = *chunk_tiles;
    const uint32_t qrows = seed_tiles ? (uint32_t)kTcSeedRows : qrows_main;
    const uint32_t nqt = (cnt + qrows - 1) / qrows;
    uint32_t ntiles = (ngl + kTcTileGroups - 1) / kTcTileGroups;
    if (seed_tiles) ntiles = min(ntiles, seed_tiles);
    const uint32_t arow0 = a_rowoff ? a_rowoff[l] : 0u;  // streamed query tiles (D > 512): first row of the list's tiles
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        // chunk-major: CTAs running at the same time share a vector chunk (L2 reuse), and a query's later chunks
        // start with a warm bound
        const uint32_t c = i / nqt, qt = i - c * nqt;
        // (chunks are NOT equalised within a list: full chunks first, the short remainder last -- chunk-major order then
        // hands out the long items first, which measured 2.5 % better than equal chunks)
        TcItem r;
        r.t0 = c * chunk;
        r.t1 = min(ntiles, r.t0 + chunk);
        r.qbase = qoff + qt * qrows;
        r.nq_tile = min(qrows, cnt - qt * qrows);
        r.g_list = g_list;
        r.ngl = ngl;
        r.valid = 1u;
        r.pad = a_rowoff ? arow0 + qt * qrows : 0u;  // all tiles of a list but the last have qrows rows
        items[i0 + i] = r;
    }
}

// Streamed query tiles (D > 512).  The rows of list l's work items are list_qlist[list_qoff[l] ..+ list_cnt[l]), cut into
// tiles of 128; tile (l, qt) is stored as [chunk of 8 dims][r8 rows][16 B] with r8 = its row count rounded up to 8 -- the
// tcgen05 K-major operand layout with LBO = r8 * 16 -- starting at row a_rowoff[l] + 128 * qt of a_tiles (a "row" = Dh chunks).
__global__ void tc_arows_kernel(const uint32_t* __restrict__ list_cnt, uint32_t nlist, uint32_t* __restrict__ rows8) {
    uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < nlist) rows8[l] = (list_cnt[l] + 7u) & ~7u;
}
// 32 consecutive entries of list_qlist per block (one per lane), the eight warps take every eighth chunk: a warp's stores are
// 512 contiguous bytes, its loads one full 32-byte sector per lane.
__global__ void __launch_bounds__(256) tc_atile_kernel(const uint32_t* __restrict__ list_cnt, const uint32_t* __restrict__ list_qoff,
                                                       const uint2* __restrict__ list_qlist, const uint32_t* __restrict__ a_rowoff,
                                                       uint32_t nlist, const float4* __restrict__ xq4, int Dq, int Dh,
                                                       const TcScale* __restrict__ scale, uint4* __restrict__ a_tiles) {
    __shared__ uint32_t s_query[32], s_tile0[32], s_row[32], s_r8[32], s_fill[32];
    const uint32_t total = list_qoff[nlist];
    if (threadIdx.x < 32) {
        const uint32_t pos = blockIdx.x * 32u + threadIdx.x;
        uint32_t q = kNoRow, tile0 = 0, row = 0, r8 = 8, fill = 0;
        if (pos < total) {
            uint32_t lo = 0, hi = nlist;  // the list with list_qoff[l] <= pos < list_qoff[l + 1] (empty lists share an offset)
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (list_qoff[mid] <= pos) lo = mid;
                else hi = mid;
            }
            const uint32_t i = pos - list_qoff[lo], cnt = list_cnt[lo], qt = i >> 7;
            const uint32_t rows = min(128u, cnt - qt * 128u);
            q = list_qlist[pos].x;
            row = i & 127u;
            r8 = (rows + 7u) & ~7u;
            tile0 = a_rowoff[lo] + qt * 128u;
            fill = i == cnt - 1 ? r8 - rows : 0u;  // the list's last row also zeroes the padding rows behind it
        }
        s_query[threadIdx.x] = q;
        s_tile0[threadIdx.x] = tile0;
        s_row[threadIdx.x] = row;
        s_r8[threadIdx.x] = r8;
        s_fill[threadIdx.x] = fill;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t q = s_query[lane];
    if (q == kNoRow) return;
    const uint32_t r8 = s_r8[lane], fill = s_fill[lane];
    const float qmul = scale->qmul;
    uint4* tile = a_tiles + (size_t)s_tile0[lane] * (size_t)Dh + s_row[lane];
    const float4* src = xq4 + (size_t)q * Dq;
    for (int c = w; c < Dh; c += 8) {
        float4 a = make_float4(0, 0, 0, 0), b = a;
        if (2 * c < Dq) a = __ldg(src + 2 * c);
        if (2 * c + 1 < Dq) b = __ldg(src + 2 * c + 1);
        uint4* dst = tile + (size_t)c * r8;
        *dst = make_uint4(pack_h2(qmul * a.x, qmul * a.y), pack_h2(qmul * a.z, qmul * a.w), pack_h2(qmul * b.x, qmul * b.y),
                          pack_h2(qmul * b.z, qmul * b.w));
        for (uint32_t z = 1; z <= fill; z++) dst[z] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ------------------------------------------------------------------------------------------
// the tensor-core scan
// ------------------------------------------------------------------------------------------
struct TcSmemLayout {
    uint32_t a_bytes, off_b, off_norm, off_ones, off_zero, off_r, off_queue, off_stage, off_q, off_row, off_bar, off_misc, off_item, total;
    uint32_t stages, stage_data, stage_bytes;
    uint32_t stream_a;  // 1: the query tile does not stay in shared memory -- its K-slices travel through the ring with the list's
};
// pair = the CTA-pair kernel: a CTA holds 64 of a tile's 128 vectors, so the same ring memory makes 8 stages instead of 4
__host__ __device__ inline TcSmemLayout tc_smem_layout_n(int Dh, int kr, bool pair, uint32_t stages) {
    TcSmemLayout L;
    L.stages = stages;
    L.stream_a = 0;
    L.stage_data = pair ? kTcStageData / 2 : kTcStageData;
    L.stage_bytes = pair ? kTcStageBytes / 2 : kTcStageBytes;
    L.a_bytes = (uint32_t)Dh * kTcM * 16;                   // query tile, [chunk][128 rows][16 B = 8 halfs]
    L.off_b = L.a_bytes;                                    // ring of list-tile K-slices
    L.off_norm = L.off_b + L.stages * L.stage_bytes;        // (end of the ring; every stage carries its own norm chunk)
    L.off_ones = L.off_norm;                                // A-side partner of the norm chunk: (a,a,a,0,..) per row
    L.off_zero = L.off_ones + 2048;                         // second K chunk of the norm step, both operands: zeros
    L.off_r = L.off_zero + 2048;                            // per query row: its k smallest filter values, descending
    L.off_queue = L.off_r + (uint32_t)kTcM * (kr + 1) * 4;  // (row stride kr+1: lanes on different rows hit different banks)
    L.off_stage = L.off_queue + tc_queue_cap(kr) * 8;            // hit queue: epilogue threads -> selector warp; then survivor staging
    L.off_q = L.off_stage + tc_stage_cap(kr) * 12;               // (query, probe rank) of the tile's rows
    L.off_row = L.off_q + kTcM * 8;                         // per row: bound P, delta, base, improved flag
    L.off_bar = L.off_row + 4 * kTcM * 4;
    L.off_misc = L.off_bar + (2 * 2 * kTcStages + 16) * 8;  // 8 + 8 ring barriers (pair kernel), 16 accumulator barriers (A-in-TMEM)
    L.off_item = L.off_misc + 64 + 512;                      // two staged work-item records
    L.total = L.off_item + 2 * 32;
    return L;
}
// The whole query tile stays in shared memory for the life of a work item (every list tile is multiplied with it), so the ring
// gets what is left: four 34 KB stages up to D = 256, two (one per tile pipeline) up to D = 512.  pair = the CTA-pair kernel:
// a CTA holds 64 of a tile's 128 vectors, so the same ring memory makes twice the stages.
// Beyond D = 512 (the reference's own grids use 768 and 1536: bench.yaml:1-15, tests/ivf_index_tests.rs:661-686) the query tile is
// STREAMED: a pre-pass writes every work item's query tile as fp16 in operand layout to global memory (tc_atile_kernel), and a
// ring stage carries 64 dimensions of the list tile AND of the query tile (stream_a).  The tile then comes from L2 once per list
// tile instead of once per work item -- half the tensor rate, but ~10 x the exact FP32 kernels these shapes used to fall back to.
constexpr uint32_t kTcSmemMax = 227 * 1024;
__host__ __device__ inline TcSmemLayout tc_smem_layout_resident(int Dh, int kr, bool pair) {
    uint32_t stages = pair ? 2 * kTcStages : kTcStages;
    TcSmemLayout L = tc_smem_layout_n(Dh, kr, pair, stages);
    while (L.total > kTcSmemMax && stages > 2) {
        stages >>= 1;
        L = tc_smem_layout_n(Dh, kr, pair, stages);
    }
    return L;
}
__host__ __device__ inline TcSmemLayout tc_smem_layout_streamed(int kr) {
    TcSmemLayout L = tc_smem_layout_n(0, kr, false, kTcStages);
    L.stream_a = 1;
    L.stage_data = kTcStreamChunks * kTcTileGroups * 512;
    return L;
}
inline TcSmemLayout tc_smem_layout(int Dh, int kr, bool pair = false) {  // host: which of the two a shape gets
    const TcSmemLayout L = tc_smem_layout_resident(Dh, kr, pair);
    return L.total > kTcSmemMax && !pair ? tc_smem_layout_streamed(kr) : L;
}

__device__ __forceinline__ uint32_t lds_volatile(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ float lds_volatile_f(const float* p) { return __uint_as_float(lds_volatile(reinterpret_cast<const uint32_t*>(p))); }
__device__ __forceinline__ void sts_volatile(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 lds_volatile_v2(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile_v2(uint2* p, uint2 v) {
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y) : "memory");
}
// min of 32 registers as a tree of 3-input minima (FMNMX3); NaNs (padding rows) drop out
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float min32(const float* v) {
    float a0 = min3(v[0], v[1], v[2]), a1 = min3(v[3], v[4], v[5]), a2 = min3(v[6], v[7], v[8]), a3 = min3(v[9], v[10], v[11]);
    float a4 = min3(v[12], v[13], v[14]), a5 = min3(v[15], v[16], v[17]), a6 = min3(v[18], v[19], v[20]);
    float a7 = min3(v[21], v[22], v[23]), a8 = min3(v[24], v[25], v[26]), a9 = min3(v[27], v[28], v[29]);
    float b0 = min3(a0, a1, a2), b1 = min3(a3, a4, a5), b2 = min3(a6, a7, a8), b3 = min3(a9, v[30], v[31]);
    return fminf(fminf(b0, b1), fminf(b2, b3));
}

// Queue entry (8 bytes, written with one 64-bit store): x = 0x40000000 | tile (10 bits) << 14 | column << 7 | row
// (never zero; zero marks an empty slot), y = float bits of the value: a filter value that passed the row's
// bound (a survivor candidate).
constexpr uint32_t kEntValid = 0x40000000u;

// Optional role timers (-DVIDX_TC_TIMING): cycles each role spends in its waits, summed per CTA into p.dbg[16 * blockIdx.x + slot].
#ifdef VIDX_TC_TIMING
#define TC_T0() const long long _t0 = clock64()
#define TC_ACC(slot) do { if (lane == 0 && p.dbg) atomicAdd(&p.dbg[16 * blockIdx.x + (slot)], (unsigned long long)(clock64() - _t0)); } while (0)
#else
#define TC_T0() do { } while (0)
#define TC_ACC(slot) do { } while (0)
#endif
// PAIR: two CTAs of a cluster (one TPC) share every tile -- tcgen05 cta_group::2, M = 256: each CTA keeps 128 query rows
// and their accumulators, loads only HALF of every list tile (64 vectors) and the tensor cores of both SMs read both halves.
// One issued instruction feeds two tensor pipes (the single-thread issue cadence, ~125 cycles per MMA, no longer bounds a
// 64-cycle MMA) and the L2 -> SM traffic per flop halves.  Only the leader CTA issues; completion is multicast to the
// barriers of both CTAs; the peer forwards "my half landed" and "my accumulator stage is drained" to the leader's barriers.
// TSA: the query tile lives in tensor memory (tcgen05.mma with A from TMEM).  The accumulators then have three stages, shared by
// the two tile pipelines in tile order (stage = tile % 3); the hand-over barriers are per pipeline and four deep (tfull[pipe][j],
// tdone[pipe][j], j = (tile / 2) % 4), so that every barrier is waited on by exactly one warp role, phase after phase -- a
// parity wait cannot tell phase n from phase n + 2, and with stage-indexed barriers shared by two issuers a slow issuer would
// read a stale "drained".
// SA: the query tile is streamed through the ring (D > 512, see tc_smem_layout): K-slices of 8 chunks, a stage = list chunks +
// norm chunk + the query tile's chunks.  A streamed tile has only as many rows as the item has queries (rounded up to 8): the
// MMA still reads 128 rows, the rows beyond are whatever the stage holds -- rows of the accumulator nobody looks at.
// NB: the norm term is added in the epilogue instead of by a ninth MMA per tile (main pass of the list scan only).  The shadow
// rows are sorted by norm inside every 1024-vector segment, so the 32 columns of a group have nearly the same norm term: a row's
// minimum over the group plus the group's SMALLEST norm term (p.gmin, one float4 per tile, fetched while the thread waits for
// the accumulator) is a lower bound of every filter value of the group, and only when that passes the row's bound -- the rare
// path, which already costs a branch -- are the columns' own norm terms (p.vnorm32) looked at.  Saves 1/9 of the tensor work
// at D = 128 and the 2 KB norm chunk per tile.
template <int KR, bool PAIR, bool TSA, bool SA, bool NB>
__global__ void __launch_bounds__(kTcThreads, 1) scan_tc_kernel(const __grid_constant__ TcParams p) {
    static_assert(!(PAIR && TSA), "the A-in-TMEM variant is single-CTA");
    static_assert(!(SA && (PAIR || TSA)), "the streamed query tile is a variant of the default kernel only");
    static_assert(!(NB && (PAIR || TSA || SA)), "epilogue-side norms are a variant of the default kernel only");
    extern __shared__ __align__(1024) unsigned char smem[];
    const TcSmemLayout L = SA ? tc_smem_layout_streamed(KR) : tc_smem_layout_resident(p.Dh, KR, PAIR);
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    constexpr uint32_t kHalfRows = PAIR ? 64u : 128u;  // vectors of a tile in this CTA's shared memory
    constexpr int kSC = SA ? kTcStreamChunks : kTcStageChunks;  // 16-byte chunks per K-slice
    constexpr uint32_t kStageData = (uint32_t)kSC * kHalfRows * 16u;  // list chunks of a stage; then the tile's norm chunk
    constexpr uint32_t kStageA = kStageData + kHalfRows * 16u;        // SA: then the query tile's chunks
    constexpr uint32_t kStageBytes = (uint32_t)kTcStageChunks * kHalfRows * 16u + kHalfRows * 16u;
    unsigned char* sA = smem;
    unsigned char* sB = smem + L.off_b;
    constexpr int kRS = KR + 1;                                     // row stride of s_r
    float* s_r = reinterpret_cast<float*>(smem + L.off_r);          // [128][KR+1]
    constexpr int kTcQueueCap = tc_queue_cap(KR) / kTcSelectors, kTcStageCap = tc_stage_cap(KR) / kTcSelectors;  // per selector
    uint2* s_stage_all = reinterpret_cast<uint2*>(smem + L.off_stage);  // [2][kTcStageCap] (row id, query row)
    float* s_stage_v_all = reinterpret_cast<float*>(smem + L.off_stage + kTcSelectors * kTcStageCap * 8);  // [2][kTcStageCap] filter value
    uint2* s_queue_all = reinterpret_cast<uint2*>(smem + L.off_queue);  // [2][kTcQueueCap]
    uint2* s_q = reinterpret_cast<uint2*>(smem + L.off_q);
    float* s_P = reinterpret_cast<float*>(smem + L.off_row);        // [128] bound in filter space (written by the selector only)
    float* s_delta = s_P + kTcM;                                    // [128]
    float* s_base = s_delta + kTcM;                                 // [128]
    uint32_t* s_impr = reinterpret_cast<uint32_t*>(s_base + kTcM);  // [128] the row's set changed during this item
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + L.off_bar);  // [stages] K-slice landed
    const uint32_t nstages = L.stages, kSP = nstages / 2;                // stages in all / per tile pipeline (1, 2 or 4)
    const uint32_t kSPs = kSP == 4 ? 2u : (kSP == 2 ? 1u : 0u);          // log2
    constexpr uint32_t kMaxStages = 2 * kTcStages;
    uint64_t* bar_empty = bar_full + kMaxStages;                         // [stages] K-slice consumed by the MMAs
    uint64_t* bar_tfull = bar_empty + kMaxStages;                        // [4] accumulator tile complete   (TSA: [2][4] per pipeline)
    uint64_t* bar_tempty = bar_tfull + (TSA ? 8 : kTcAccStages);         // [4] accumulator tile drained    (TSA: [2][4] "group drained its tile")
    uint32_t* s_misc = reinterpret_cast<uint32_t*>(smem + L.off_misc);   // [0] tmem base [1] item [2] queue tail [3] queue head [4] done
    volatile float* s_tmpv_all = reinterpret_cast<volatile float*>(s_misc + 16);  // [4][32] selector scratch
    // s_misc: [0] tmem base, [4] epilogue warps done, [8 + 2*sel] queue tail, [9 + 2*sel] queue head (sel < 4)

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(kFull, tid >> 5, 0);  // warp-uniform for the compiler too
    const int Dq = p.Dq, Dh = p.Dh;
    // the batch's scales (uniform loads); a batch that cannot be scaled into fp16 range goes to the exact kernels
    const float tS = p.scale->S, tInvS = p.scale->invS, tQmul = p.scale->qmul, tCabs = p.scale->c_abs;
    const float tAones = NB ? p.scale->a_ones : 0.0f;  // 2^(sq-sv+g): the stored norm terms times this are in accumulator units
    // flags bit 0: the rows' k-smallest sets stay CTA-local -- gtop (written by the bounds pass) is read-only in the main
    // pass, no seqlock / lock traffic at item boundaries; CTAs still share their bounds through gthr (atomicMin)
    const bool local_sets = (p.flags & 1u) != 0;
    if (!p.scale->ok) {
        for (uint32_t q = blockIdx.x * kTcThreads + tid; q < p.nq; q += gridDim.x * kTcThreads) p.overflow[q] = 1u;
        return;
    }
    if (tid == 0) {
        for (uint32_t i = 0; i < nstages; i++) {
            mbar_init(&bar_full[i], PAIR && cta_rank == 0 ? 2 : 1);  // leader of a pair: its own copies + the peer's "landed"
            mbar_init(&bar_empty[i], 1);
        }
        for (int i = 0; i < (TSA ? 8 : kTcAccStages); i++) {
            mbar_init(&bar_tfull[i], 1);
            // the four warps of the epilogue group that owns the tile (pair: of both CTAs, all on the leader's barrier)
            mbar_init(&bar_tempty[i], PAIR ? kTcEpiWarps : kTcEpiWarps / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // constant operands of the norm step and an empty queue
    for (int i = tid; i < 128; i += kTcThreads) {
        const float a = p.scale->a_ones;  // (a, a, a, 0, 0, 0, 0, 0) per row, as fp16
        reinterpret_cast<uint4*>(smem + L.off_ones)[i] = make_uint4(pack_h2(a, a), pack_h2(a, 0.0f), 0u, 0u);
        reinterpret_cast<uint4*>(smem + L.off_zero)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < kTcSelectors * kTcQueueCap; i += kTcThreads) s_queue_all[i] = make_uint2(0u, 0u);
    if (warp == 1) {
        if (PAIR) {  // collective over the pair: one warp of each CTA
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_misc[0])), "r"(kTcTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_misc[0])), "r"(kTcTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_misc[0];
    if (TSA && warp >= 5 && warp <= 8) {
        // the A side of the norm step, (a, a, a, 0, ...) per row as fp16, once: 8 columns behind the query tile
        const float a = p.scale->a_ones;
        const uint32_t ones[8] = {pack_h2(a, a), pack_h2(a, 0.0f), 0u, 0u, 0u, 0u, 0u, 0u};
        tc_st8(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTcACol + (uint32_t)p.Dh * 4, ones);
        tc_wait_st();
        tc_fence_before();
    }
    const uint32_t total_items = p.item_off[p.nlist];
    const uint32_t idesc = make_idesc_f16(PAIR ? 2 * kTcM : kTcM, kTcTileGroups * 32);
    const int nkc = (Dh + kSC - 1) / kSC;  // K-slices per tile
    const float kInf = __int_as_float(0x7f800000);
    uint32_t it = 0;     // tiles processed so far by this CTA (accumulator stage = it & 3, phase = (it >> 2) & 1)
    // Two independent tile pipelines, each with its own producer warp, MMA warp and half of the ring: pipeline 0
    // takes the even tiles of this CTA's tile sequence, pipeline 1 the odd ones.  (One ring shared by two issuers
    // would let an issuer run a whole ring ahead of the other, where a phase-parity wait reads a stale "ready".)
    uint32_t ks_it = 0;  // K-slices processed so far by this warp's pipeline (stage = 2*pipe + (ks_it & 1), phase = (ks_it >> 1) & 1)
    uint32_t ks_it1 = 0; // (producer: the same for pipeline 1; ks_it is pipeline 0's)

    // Work items are claimed one ahead: while item i runs, the selector warp (idle during an item's set-up) claims
    // item i+1 and stages its record in shared memory.
    TcItem* s_item = reinterpret_cast<TcItem*>(smem + L.off_item);  // [2]
    auto claim = [&](int slot) {  // one thread (of the leader CTA: both CTAs of a pair work on the same item)
        const uint32_t idx = atomicAdd(p.work_counter, 1u);
        TcItem r;
        if (idx < total_items) r = p.items[idx];
        else r.valid = 0u;
        s_item[slot] = r;
        if (PAIR) {
            const uint32_t peer = mapa_u32(smem_u32(&s_item[slot]), 1u);
            st_cluster_v4(peer, make_uint4(r.t0, r.t1, r.qbase, r.nq_tile));
            st_cluster_v4(peer + 16u, make_uint4(r.g_list, r.ngl, r.valid, 0u));
        }
    };
    if (tid == 0 && cta_rank == 0) claim(0);
    uint32_t cur = 0;
#ifdef VIDX_TC_TIMING
    const long long _tk0 = clock64();
#endif
    for (;; cur ^= 1u) {
#ifdef VIDX_TC_TIMING
        const long long _ti0 = clock64();
#endif
        // the record of this item is staged (pair: in both CTAs); every warp is done with the previous item
        if (PAIR) cluster_sync_all();
        else __syncthreads();
#ifdef VIDX_TC_TIMING
        if (tid == 64 && p.dbg) atomicAdd(&p.dbg[16 * blockIdx.x + 9], (unsigned long long)(clock64() - _ti0));  // warp 2: drain wait
#endif
        const TcItem rec = s_item[cur];
        if (!__shfl_sync(kFull, rec.valid, 0)) break;
        const uint32_t t0 = __shfl_sync(kFull, rec.t0, 0), t1 = __shfl_sync(kFull, rec.t1, 0);
        const uint32_t g_list = __shfl_sync(kFull, rec.g_list, 0), ngl = __shfl_sync(kFull, rec.ngl, 0);
        const uint32_t nq_tile = __shfl_sync(kFull, rec.nq_tile, 0), qbase = __shfl_sync(kFull, rec.qbase, 0);
        // SA: the item's query tile in p.a_tiles: first row, and its row count (a multiple of 8: whole core matrices)
        const uint32_t a_row0 = SA ? __shfl_sync(kFull, rec.pad, 0) : 0u;
        const uint32_t a_r8 = SA ? ((min(nq_tile, (uint32_t)kTcM) + 7u) & ~7u) : (uint32_t)kTcM;

        // The producers start streaming this item's tiles at once (they need nothing but the record); everyone else
        // sets the item up behind two named barriers the producers do not take part in.
        if (warp != 0) {
            constexpr int kSetupThreads = kTcThreads - 32;
            if (tid == 352) {  // (a selector warp) empty queues for this item
                s_misc[4] = 0;
                for (int i = 8; i < 16; i++) s_misc[i] = 0;  // queue tail / head
            }
            const int row = tid - 32;  // warps 1-4 own the 128 query rows during set-up
            uint2 qi = make_uint2(kNoRow, 0);
            if (row >= 0 && row < kTcM) {
                const uint32_t irow = cta_rank * kTcM + (uint32_t)row;  // row of the work item (pair: 256 rows, 128 per CTA)
                if (irow < nq_tile) qi = p.list_qlist[qbase + irow];
                s_q[row] = qi;
            }
            asm volatile("bar.sync 2, %0;" ::"n"(kSetupThreads) : "memory");
            if (row >= 0 && row < kTcM) {
                // per-row state of this item: its top-k set as known to all CTAs, its bound
                float P = -kInf, delta = 0.0f, base_t = 0.0f;
                float* rr = s_r + row * kRS;
                uint32_t dump_off = 0;
                if (p.mode == 2) {
                    // bounds pass: no bounds yet; the row's sub-tile minima go to its (query, probe rank) row of submin
                    if (qi.x != kNoRow) dump_off = p.pair_off ? p.pair_off[(size_t)qi.x * p.nprobe + qi.y] : (qi.x * p.seed_ranks + qi.y) * p.noinsert_tiles;
                } else if (qi.x != kNoRow) {
                    const uint32_t q = qi.x;
                    const float qn = p.qnorm[q];
                    // base_t in real units; delta, the set and P in accumulator units (x S, a power of two)
                    base_t = (1.0f - kTcEps) * qn - tCabs;
                    delta = (2.0f * kTcEps * (qn + p.vn_max) + 2.0f * tCabs) * tS;
                    const float g = __uint_as_float(__ldcg(&p.gthr_bits[q]));
                    const float tau_g = ((g - base_t) + 1e-5f * (g + fabsf(base_t))) * tS;
                    // seqlock read: writers make the version odd while they update the set
                    const volatile uint32_t* ver = p.gver + q;
                    float r0 = kInf;
                    if (local_sets) {
                        for (int i = 0; i < KR; i++) {
                            float v = i < (int)p.k ? __ldcg(&p.gtop[(size_t)q * p.k + i]) : -kInf;
                            rr[i] = v;
                            if (i == 0) r0 = v;
                        }
                    } else
                    for (;;) {
                        uint32_t v1 = *ver;
                        if (v1 & 1u) { __nanosleep(32); continue; }
                        __threadfence();
                        for (int i = 0; i < KR; i++) {
                            float v = i < (int)p.k ? __ldcg(&p.gtop[(size_t)q * p.k + i]) : -kInf;
                            rr[i] = v;
                            if (i == 0) r0 = v;
                        }
                        __threadfence();
                        if (*ver == v1) break;
                    }
                    P = fminf(tau_g, r0 + delta);
                }
                s_P[row] = P;
                s_delta[row] = delta;
                s_base[row] = base_t;
                s_impr[row] = dump_off;  // (bounds pass: first tile of the row in submin; else "set changed" flag = 0)
            } else if (TSA) {
                // A tile = fp16(-2 * 2^sq * queries) in tensor memory: row r in lane r, 8 dimensions per 4 columns; dimensions
                // beyond the query's are zero.  A warp reaches the TMEM lanes of its quarter (warp % 4) only: warps 5-8.
                if (warp >= 5 && warp <= 8) {
                    const int arow = (warp & 3) * 32 + lane;
                    const uint32_t q = s_q[arow].x;
                    const float4* src = p.xq4 + (size_t)(q == kNoRow ? 0u : q) * Dq;
                    const uint32_t abase = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTcACol;
                    for (int ks = 0; ks < Dh / 2; ks += 2) {  // two K steps (32 dimensions, 8 loads) in flight
                        float4 v[8];
#pragma unroll
                        for (int u8 = 0; u8 < 8; u8++) {
                            const int c4 = ks * 4 + u8;
                            v[u8] = (q != kNoRow && c4 < Dq) ? __ldg(src + c4) : make_float4(0, 0, 0, 0);
                        }
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            if (ks + h < Dh / 2) {
                                uint32_t r8[8];
#pragma unroll
                                for (int u4 = 0; u4 < 4; u4++) {
                                    const float4 x = v[4 * h + u4];
                                    r8[2 * u4] = pack_h2(tQmul * x.x, tQmul * x.y);
                                    r8[2 * u4 + 1] = pack_h2(tQmul * x.z, tQmul * x.w);
                                }
                                tc_st8(abase + (uint32_t)(ks + h) * 8, r8);
                            }
                        }
                    }
                    tc_wait_st();
                    tc_fence_before();
                }
            } else if (!SA && warp >= 5) {
                // A tile = fp16(-2 * 2^sq * queries): [chunk of 8 dims][128 rows][16 B] (core matrices of 8 rows x
                // 16 B, SBO 128 B, LBO 2048 B); dimensions beyond the query's are zero.  Warps 5-11 gather it while
                // warps 1-4 fetch the row state.
                constexpr int kGatherThreads = 7 * 32;  // warps 5-11
                const int gt = tid - 160;
                for (int base = 0; base < Dh * kTcM; base += kGatherThreads * 4) {
                    float4 va[4], vb[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {  // issue the gathers first, then store: 8 loads in flight per thread
                        const int idx = base + u * kGatherThreads + gt;
                        va[u] = make_float4(0, 0, 0, 0);
                        vb[u] = va[u];
                        if (idx < Dh * kTcM) {
                            const uint32_t q = s_q[idx & 127].x;
                            const int c = idx >> 7;
                            if (q != kNoRow) {
                                if (2 * c < Dq) va[u] = __ldg(&p.xq4[(size_t)q * Dq + 2 * c]);
                                if (2 * c + 1 < Dq) vb[u] = __ldg(&p.xq4[(size_t)q * Dq + 2 * c + 1]);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int idx = base + u * kGatherThreads + gt;
                        if (idx < Dh * kTcM)
                            reinterpret_cast<uint4*>(sA)[idx] =
                                make_uint4(pack_h2(tQmul * va[u].x, tQmul * va[u].y), pack_h2(tQmul * va[u].z, tQmul * va[u].w),
                                           pack_h2(tQmul * vb[u].x, tQmul * vb[u].y), pack_h2(tQmul * vb[u].z, tQmul * vb[u].w));
                    }
                }
                fence_proxy_async();
            }
            asm volatile("bar.sync 2, %0;" ::"n"(kSetupThreads) : "memory");
        }
#ifdef VIDX_TC_TIMING
        if (tid == 64 && p.dbg) {
            atomicAdd(&p.dbg[16 * blockIdx.x + 10], (unsigned long long)(clock64() - _ti0));  // loop top -> roles start
            atomicAdd(&p.dbg[16 * blockIdx.x + 11], 1ull);                                    // items
            atomicAdd(&p.dbg[16 * blockIdx.x + 12], (unsigned long long)(t1 - t0));           // tiles
        }
#endif

        if (warp == 0) {
            // ===== producer: feeds both pipelines' halves of the ring in tile order.  The warp runs the loop converged
            // (warp-uniform values), one elected lane issues.  A list chunk is one linear stream in HBM (tiles and
            // their K-slices are consecutive), so the source just advances. =====
            const uint32_t tile_bytes = (uint32_t)Dh * kSuper * 16;
            const unsigned char* src0 = reinterpret_cast<const unsigned char*>(p.vecs16) + ((size_t)(g_list >> 2) + t0) * tile_bytes;
            const uint4* nsrc0 = p.vnorm + ((size_t)g_list + (size_t)t0 * kTcTileGroups) * 32;
            const unsigned char* asrc0 = SA ? reinterpret_cast<const unsigned char*>(p.a_tiles) + (size_t)a_row0 * (size_t)Dh * 16 : nullptr;
            // One K-slice of tile t into ring stage s.
            auto load_slice = [&](uint32_t t, int kc, uint32_t s) {
                const uint32_t nch = (uint32_t)min(kSC, Dh - kc * kSC);
                const uint32_t bytes = nch * kSuper * 16;  // of the whole 128-vector K-slice in HBM
                const unsigned char* src = src0 + (size_t)(t - t0) * tile_bytes + (size_t)kc * (kSC * kSuper * 16);
                const uint4* nsrc = nsrc0 + (size_t)(t - t0) * kSuper;
                const bool last = kc == nkc - 1;
                if (elect_one()) {
                    if (PAIR) {
                        // this CTA's 64 vectors of every 16-byte chunk -- 1 KB out of each 2 KB chunk block -- as ONE tiled
                        // copy: the shadow store seen as [chunk blocks][2 KB], box = 16 blocks x 1 KB (sixteen 1 KB bulk copies
                        // per K-slice made the producer the slowest role).  The box always has 16 rows: a short last slice
                        // pulls in chunks of the next tile, which no MMA reads.
                        const uint32_t blk0 = ((g_list >> 2) + t) * (uint32_t)Dh;  // first 2 KB chunk block of the tile
                        mbar_expect_tx(&bar_full[s], kStageData + (last ? 1024u : 0u));
                        tma_load_2d(sB + s * kStageBytes, &p.tmap, (int)(cta_rank * 128u), (int)(blk0 + (uint32_t)kc * kTcStageChunks), &bar_full[s]);
                        if (last) bulk_g2s(sB + s * kStageBytes + kStageData, nsrc + cta_rank * 64u, 1024u, &bar_full[s]);
                    } else {
                        const uint32_t abytes = SA ? nch * a_r8 * 16u : 0u;  // the same K-slice of the query tile: [chunk][a_r8 rows][16 B]
                        const bool nchunk = last && !NB;                     // the tile's norm chunk rides with its last K-slice
                        mbar_expect_tx(&bar_full[s], bytes + (nchunk ? 2048u : 0u) + abytes);
                        bulk_g2s(sB + s * kStageBytes, src, bytes, &bar_full[s]);
                        if (nchunk) bulk_g2s(sB + s * kStageBytes + kStageData, nsrc, 2048, &bar_full[s]);
                        if (SA) bulk_g2s(sB + s * kStageBytes + kStageA, asrc0 + (size_t)kc * ((size_t)kSC * a_r8 * 16u), abytes, &bar_full[s]);
                    }
                }
                __syncwarp();
            };
            if (!(p.flags & 8u)) {
                // strict tile order (default): a pipeline whose ring half is full stalls the other one too -- measured equal to
                // the independent feeding below (2.089 vs 2.095 ms on the bench workload), so the simpler order stays
                for (uint32_t t = t0; t < t1; t++) {
                    const uint32_t pipe = (it + (t - t0)) & 1u;
                    for (int kc = 0; kc < nkc; kc++) {
                        const uint32_t cnt = pipe ? ks_it1++ : ks_it++;
                        const uint32_t s = kSP * pipe + (cnt & (kSP - 1u)), ph = (cnt >> kSPs) & 1;
                        { TC_T0(); mbar_wait(&bar_empty[s], ph ^ 1); TC_ACC(0 + pipe); }
                        load_slice(t, kc, s);
                    }
                }
            } else {
                // (VIDX_TC_FLAGS bit 3) each pipeline fed independently: a cursor per pipeline (its tiles are every other tile of the
                // item), the producer serves whichever ring half has a free stage, so that a pipeline whose epilogue falls behind
                // cannot starve the other one through the producer.
                uint32_t tcur[2] = {t0 + ((0u ^ it) & 1u), t0 + ((1u ^ it) & 1u)};  // first tile of pipeline 0 / 1 in this item
                int kcur[2] = {0, 0};
                uint32_t idle = 0;
                while (tcur[0] < t1 || tcur[1] < t1) {
                    bool progressed = false;
#pragma unroll
                    for (uint32_t pipe = 0; pipe < 2; pipe++) {
                        if (tcur[pipe] >= t1) continue;
                        const uint32_t cnt = pipe ? ks_it1 : ks_it;
                        const uint32_t s = kSP * pipe + (cnt & (kSP - 1u)), ph = (cnt >> kSPs) & 1;
                        // (one lane decides: the attempts of different lanes may see different phases)
                        if (!__shfl_sync(kFull, mbar_try_wait(&bar_empty[s], ph ^ 1) ? 1 : 0, 0)) continue;
                        load_slice(tcur[pipe], kcur[pipe], s);
                        if (pipe) ks_it1++; else ks_it++;
                        if (++kcur[pipe] == nkc) {
                            kcur[pipe] = 0;
                            tcur[pipe] += 2;
                        }
                        progressed = true;
                    }
                    if (progressed) idle = 0;
                    else if (++idle > (1u << 24)) __trap();  // a protocol bug, do not hang the GPU
                }
            }
            it += t1 - t0;
        } else if (warp == 9 || warp == 10) {
            // ===== MMA issuers: warp 1 takes the even tiles of this CTA's tile sequence, warp 10 the odd ones
            // (one thread each; the loop is kept to a few dozen instructions per K-slice because a single
            // thread's issue latency, not the tensor pipe, would otherwise bound the kernel) =====
            const uint32_t pipe = warp == 9 ? 0u : 1u;
            if (PAIR && cta_rank != 0) {
                // the peer CTA issues nothing: this warp tells the leader when this CTA's half of a K-slice has landed
                const uint32_t skip = (pipe ^ it) & 1u;
                for (uint32_t itt = it + skip; itt < it + (t1 - t0); itt += 2) {
                    for (int kc = 0; kc < nkc; kc++, ks_it++) {
                        const uint32_t s = kSP * pipe + (ks_it & (kSP - 1u)), ph = (ks_it >> kSPs) & 1;
                        mbar_wait(&bar_full[s], ph);
                        if (elect_one()) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_full[s]), 0u));
                        __syncwarp();
                    }
                }
                it += t1 - t0;
            } else {
                // descriptor = lo | hi << 32; lo = start address >> 4 (14 bits) | LBO >> 4 << 16, hi = SBO >> 4 | version 1 << 14
                const uint32_t desc_hi = (128u >> 4) | (1u << 14);
                const uint32_t lbo_bits = (2048u >> 4) << 16;
                const uint32_t a_lo0 = ((smem_u32(sA) >> 4) & 0x3fffu) | lbo_bits;
                // B: [chunk][rows in this CTA][16 B] -- consecutive K chunks are kHalfRows * 16 bytes apart
                const uint32_t b_lo0 = ((smem_u32(sB) >> 4) & 0x3fffu) | (((kHalfRows * 16u) >> 4) << 16);
                // norm step: K chunk 0 = (1,1,1,0) x (n_hi, n_mid, n_lo, 0), K chunk 1 = the shared zero block
                const uint32_t ones_lo = ((smem_u32(smem + L.off_ones) >> 4) & 0x3fffu) | (((L.off_zero - L.off_ones) >> 4) << 16);
                const uint32_t skip = (pipe ^ it) & 1u;  // first tile of this item that belongs to this pipeline
                for (uint32_t itt = it + skip; itt < it + (t1 - t0); itt += 2) {
                    uint32_t a, tf;  // accumulator stage of the tile, and the barrier its completion is signalled on
                    if (TSA) {
                        a = itt % 3u;
                        const uint32_t kk = itt >> 1;  // this pipeline's kk-th tile
                        tf = pipe * 4u + (kk & 3u);
                        if (itt >= 3) {
                            // the stage was last used by tile itt - 3, a tile of the OTHER pipeline: wait until its epilogue group
                            // has drained it (that group's ((itt - 3) / 2)-th tile)
                            const uint32_t ko = (itt - 3) >> 1;
                            TC_T0();
                            mbar_wait(&bar_tempty[(1u - pipe) * 4u + (ko & 3u)], (ko >> 2) & 1);
                            TC_ACC(4 + pipe);
                        }
                    } else {
                        a = itt & (kTcAccStages - 1);
                        tf = a;
                        const uint32_t aph = (itt / kTcAccStages) & 1;
                        TC_T0();
                        mbar_wait(&bar_tempty[a], aph ^ 1);
                        TC_ACC(4 + pipe);
                    }
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + a * 128;
                    for (int kc = 0; kc < nkc; kc++, ks_it++) {
                        const uint32_t s = kSP * pipe + (ks_it & (kSP - 1u)), ph = (ks_it >> kSPs) & 1;
                        { TC_T0(); mbar_wait(&bar_full[s], ph); TC_ACC(6 + pipe); }
                        tc_fence_after();
                        // chunk c of all 128 rows is one 2 KB block (128 x 16 B) in the query tile: +128 per chunk in >>4 units;
                        // a K step (two chunks) of the list tile is 2 * kHalfRows * 16 bytes
                        // (SA: the query tile's K-slice sits in the stage, consecutive chunks a_r8 * 16 bytes apart)
                        const uint32_t al = SA ? ((((smem_u32(sB) + s * kStageBytes + kStageA) >> 4) & 0x3fffu) | (((a_r8 * 16u) >> 4) << 16))
                                               : a_lo0 + (uint32_t)kc * (kSC * 128);
                        const uint32_t a_kstep = SA ? 2u * a_r8 : 256u;  // one K step = two chunks, in 16-byte units
                        const uint32_t bl = b_lo0 + s * (kStageBytes >> 4);
                        constexpr uint32_t kBStep = (2u * kHalfRows * 16u) >> 4;
                        const int nks = min(kSC / 2, (Dh >> 1) - kc * (kSC / 2));
                        const uint32_t noff = L.off_b + s * kStageBytes + kStageData;
                        const uint32_t norm_lo = ((smem_u32(smem + noff) >> 4) & 0x3fffu) | (((L.off_zero - noff) >> 4) << 16);
                        if (elect_one()) {
                            if (p.flags & 4u) {  // ablation (timing only): no tensor work, the ring just turns over
                                if (PAIR) { tc_commit_pair(&bar_empty[s]); if (kc == nkc - 1) tc_commit_pair(&bar_tfull[tf]); }
                                else { tc_commit(&bar_empty[s]); if (kc == nkc - 1) tc_commit(&bar_tfull[tf]); }
                            } else
                            if (PAIR) {
                                if (kc == 0) tc_mma_f16_pair<false>(d_tmem, al, bl, desc_hi, idesc);
                                else tc_mma_f16_pair<true>(d_tmem, al, bl, desc_hi, idesc);
#pragma unroll
                                for (int ks = 1; ks < kSC / 2; ks++)
                                    if (ks < nks) tc_mma_f16_pair<true>(d_tmem, al + ks * 256, bl + ks * kBStep, desc_hi, idesc);
                                if (kc == nkc - 1) tc_mma_f16_pair<true>(d_tmem, ones_lo, norm_lo, desc_hi, idesc);
                                tc_commit_pair(&bar_empty[s]);                       // both CTAs' halves of the K-slice are free
                                if (kc == nkc - 1) tc_commit_pair(&bar_tfull[tf]);   // both CTAs' accumulator tiles are ready
                            } else if (TSA) {
                                const uint32_t at = tmem_base + kTcACol + (uint32_t)kc * (kSC / 2) * 8u;  // 8 columns per K step
                                if (kc == 0) tc_mma_f16_ts<false>(d_tmem, at, bl, desc_hi, idesc);
                                else tc_mma_f16_ts<true>(d_tmem, at, bl, desc_hi, idesc);
#pragma unroll
                                for (int ks = 1; ks < kSC / 2; ks++)
                                    if (ks < nks) tc_mma_f16_ts<true>(d_tmem, at + ks * 8, bl + ks * kBStep, desc_hi, idesc);
                                if (kc == nkc - 1) tc_mma_f16_ts<true>(d_tmem, tmem_base + kTcACol + (uint32_t)Dh * 4u, norm_lo, desc_hi, idesc);
                                tc_commit(&bar_empty[s]);
                                if (kc == nkc - 1) tc_commit(&bar_tfull[tf]);
                            } else {
                                if (kc == 0) tc_mma_f16_lo<false>(d_tmem, al, bl, desc_hi, idesc);
                                else tc_mma_f16_lo<true>(d_tmem, al, bl, desc_hi, idesc);
#pragma unroll
                                for (int ks = 1; ks < kSC / 2; ks++)
                                    if (ks < nks) tc_mma_f16_lo<true>(d_tmem, al + ks * a_kstep, bl + ks * kBStep, desc_hi, idesc);
                                if (kc == nkc - 1 && !NB) tc_mma_f16_lo<true>(d_tmem, ones_lo, norm_lo, desc_hi, idesc);
                                tc_commit(&bar_empty[s]);                       // K-slice free once these MMAs have read it
                                if (kc == nkc - 1) tc_commit(&bar_tfull[tf]);   // accumulator tile ready for the epilogue
                            }
                        }
                        __syncwarp();
                    }
                }
                it += t1 - t0;
            }
        } else if (warp >= 11) {
            const uint32_t sel = (uint32_t)(warp - 11);
            uint2* s_queue = s_queue_all + sel * kTcQueueCap;
            uint2* s_stage = s_stage_all + sel * kTcStageCap;
            float* s_stage_v = s_stage_v_all + sel * kTcStageCap;
            volatile float* s_tmpv = s_tmpv_all + sel * 32;
            uint32_t* q_tail = &s_misc[8 + 2 * sel];
            uint32_t* q_head = &s_misc[9 + 2 * sel];
            // ===== selector: the only writer of the rows' bounds and top-k sets =====
            if (lane == 0 && sel == 0 && cta_rank == 0) claim((int)(cur ^ 1u));  // the next work item, one ahead
            __syncwarp();
            const uint32_t row0_item = (g_list + t0 * kTcTileGroups) * 32u;
            volatile float* vP = s_P;
            volatile float* vr = s_r;
            uint32_t head = 0, refresh = 0, idle = 0, nstage = 0;
            const bool frozen = p.frozen != 0;
            const uint32_t noinsert_tiles = p.noinsert_tiles, seed_ranks = p.seed_ranks;
            // survivors are staged in shared memory and appended to the per-query lists in bulk: four global
            // atomics in flight per lane instead of one round trip per batch; entries that fell outside the
            // row's bound meanwhile are dropped
            auto flush = [&]() {
                for (uint32_t base = 0; base < nstage; base += 128) {
                    uint32_t gi[4], rid[4];
                    uint2 qq[4];
                    bool ok[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t i = base + u * 32 + lane;
                        ok[u] = false;
                        if (i < nstage) {
                            const uint2 se = s_stage[i];
                            const uint32_t srow = se.y & 127u;
                            qq[u] = s_q[srow];
                            rid[u] = se.x;
                            ok[u] = s_stage_v[i] <= vP[srow];
                            if (ok[u] && p.perm) rid[u] = __ldg(&p.perm[se.x]);  // shadow row -> row of the fp32 store
                            if (ok[u]) gi[u] = atomicAdd(&p.cand_cnt[qq[u].x], 1u);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (ok[u]) {
                            if (gi[u] < p.capq) {
                                p.cand[(size_t)qq[u].x * p.capq + gi[u]] = ((unsigned long long)qq[u].y << 32) | rid[u];
                                if (p.cand_val) p.cand_val[(size_t)qq[u].x * p.capq + gi[u]] = s_stage_v[base + u * 32 + lane];
                            } else {
                                p.overflow[qq[u].x] = 1u;
                            }
                        }
                    }
                }
                nstage = 0;
                __syncwarp();
            };
            auto adopt = [&](int rr) {  // bounds other CTAs published for this row's query
                const uint32_t q = s_q[rr].x;
                if (q != kNoRow) {
                    const float g = __uint_as_float(__ldcg(&p.gthr_bits[q]));
                    const float b = s_base[rr];
                    const float tau = ((g - b) + 1e-5f * (g + fabsf(b))) * tS;
                    if (tau < vP[rr]) vP[rr] = tau;
                }
            };
            for (;;) {
                const uint2 e = lds_volatile_v2(&s_queue[(head + lane) & (kTcQueueCap - 1)]);
                const unsigned have = __ballot_sync(kFull, e.x != 0u);
                const int n = have == kFull ? 32 : __ffs(~have) - 1;  // longest prefix of written slots
                if (n == 0) {
                    if (lds_volatile(&s_misc[4]) == (uint32_t)kTcEpiWarps && lds_volatile(q_tail) == head) break;
                    if (++idle > (1u << 21)) __trap();  // an item never takes this long
#ifdef VIDX_TC_ABLATE
                    if (p.flags & 0x20000u) __nanosleep(256);  // (valid answers) an idle selector leaves the issue slots alone
                    if ((p.flags & 0x40000u) && (idle & 7u)) continue;  // (valid answers) ... and adopts published bounds every eighth poll only
#endif
                    for (int a4 = 0; a4 < 4; a4++) adopt(a4 * 32 + lane);
                    continue;
                }
                idle = 0;
                const bool mine = lane < n;
                if (mine) sts_volatile_v2(&s_queue[(head + lane) & (kTcQueueCap - 1)], make_uint2(0u, 0u));
                head += (uint32_t)n;
                if (lane == 0) sts_volatile(q_head, head);
                const uint32_t row = e.x & 127u;
                const float val = __uint_as_float(e.y);
#ifdef VIDX_TC_ABLATE
                if (p.flags & 0x8000u) continue;  // timing only: the selector drains the queue and drops everything
#endif
                // (1) survivors: everything still inside the row's bound goes to the exact re-check
                const bool cand_ok = mine && val <= vP[row];
                const unsigned cm = __ballot_sync(kFull, cand_ok);
                if (cand_ok) {
                    const uint32_t pos = nstage + __popc(cm & ((1u << lane) - 1u));
                    s_stage[pos] = make_uint2(row0_item + ((e.x >> 14) & 1023u) * 128u + ((e.x >> 7) & 127u), row);
                    s_stage_v[pos] = val;
                }
                nstage += __popc(cm);
                __syncwarp();
                if (nstage > (uint32_t)kTcStageCap - 32u) flush();
                // (2) state changes: one lane per distinct row folds ALL of that row's values of this batch into the
                // row's set (threads queue their hits back to back, so a batch usually holds runs of one row)
                // (values of the seeded tiles are survivors but never enter the set: the set was initialised from them)
                const bool seeded = s_q[row].y < seed_ranks && t0 + ((e.x >> 14) & 1023u) < noinsert_tiles;
                const bool todo = cand_ok && !frozen && !seeded && val < vr[row * kRS];
                const unsigned tm = __ballot_sync(kFull, todo);
                if (tm) {
                    s_tmpv[lane] = val;
                    __syncwarp();
                    if (todo) {
                        const unsigned peers = __match_any_sync(tm, row);
                        if (lane == __ffs(peers) - 1) {
                            float g[KR];
#pragma unroll
                            for (int i = 0; i < KR; i++) g[i] = vr[row * kRS + i];
                            bool ins = false;
                            for (unsigned pm = peers; pm; pm &= pm - 1) {
                                const float v = s_tmpv[__ffs(pm) - 1];
                                if (v < g[0]) {
                                    g[0] = v;  // drop the largest, bubble the new value into place (descending)
#pragma unroll
                                    for (int i = 0; i + 1 < KR; i++) {
                                        const float hi_v = fmaxf(g[i], g[i + 1]), lo_v = fminf(g[i], g[i + 1]);
                                        g[i] = hi_v;
                                        g[i + 1] = lo_v;
                                    }
                                    ins = true;
                                }
                            }
                            if (ins) {
#pragma unroll
                                for (int i = 0; i < KR; i++) vr[row * kRS + i] = g[i];
                                s_impr[row] = 1u;
                                const float newP = g[0] + s_delta[row];
                                if (newP < vP[row]) {
                                    vP[row] = newP;
                                    float U = fmaxf(newP * tInvS + s_base[row], 0.0f);
                                    U = U + 1e-5f * U;
                                    atomicMin(&p.gthr_bits[s_q[row].x], __float_as_uint(U));
                                }
                            }
                        }
                    }
                    __syncwarp();
                }
                if ((++refresh & 15u) == 0u)
                    for (int a4 = 0; a4 < 4; a4++) adopt(a4 * 32 + lane);
            }
            flush();
            it += t1 - t0;
            asm volatile("bar.sync 1, %0;" ::"n"((kTcEpiWarps + kTcSelectors) * 32) : "memory");
        } else {
            // ===== epilogue: two groups of four warps alternate tiles (group g = pipeline g's tiles); a thread owns one
            // query row of its group's tiles.  The accumulators already hold (1-eps)|v|^2 - 2 q.v (scaled), so a tile costs
            // one 3-input min per three columns and one branch per 32 =====
            const int quarter = warp & 3;        // TMEM lanes this warp may read: 32*quarter .. +31
            const uint32_t grp = (uint32_t)(warp - 1) >> 2;
            const int row = quarter * 32 + lane;
            const uint2 qi = s_q[row];
            const bool valid = qi.x != kNoRow;
            const float delta = s_delta[row];
            const uint32_t kk = p.k;
            const uint32_t submin_row0 = s_impr[row];  // bounds pass only
            float P = s_P[row];
            uint32_t nhit = 0;  // values this thread queued in this item
            float lr[KR];  // the k smallest values this thread queued in this item (from the third on), descending (+inf until k exist)
#pragma unroll
            for (int i = 0; i < KR; i++) lr[i] = i < (int)kk ? kInf : -kInf;
            const uint32_t sel = 0;  // (one selector)
            uint2* s_queue = s_queue_all + sel * kTcQueueCap;
            uint32_t* q_tail = &s_misc[8 + 2 * sel];
            uint32_t* q_head = &s_misc[9 + 2 * sel];
            auto push = [&](uint32_t info, float v) {
#ifdef VIDX_TC_ABLATE
                if (p.flags & 0x4000u) return;  // timing only: the hit path runs, nothing is queued
#endif
                const uint32_t idx = atomicAdd(q_tail, 1u);
#ifdef VIDX_TC_TIMING
                const long long _tp0 = clock64();
                bool _spun = false;
#endif
                for (uint32_t spins = 0; idx - lds_volatile(q_head) >= (uint32_t)kTcQueueCap; spins++) {
                    __nanosleep(32);
#ifdef VIDX_TC_TIMING
                    _spun = true;
#endif
                    if (spins > (1u << 22)) __trap();  // the selector always drains: a protocol bug, do not hang the GPU
                }
#ifdef VIDX_TC_TIMING
                if (p.dbg) {
                    if (_spun) atomicAdd(&p.dbg[16 * blockIdx.x + 2], (unsigned long long)(clock64() - _tp0));  // thread-cycles waiting for queue room
                    atomicAdd(&p.dbg[16 * blockIdx.x + 3], 1ull);                                                  // values queued
                }
#endif
                sts_volatile_v2(&s_queue[idx & (kTcQueueCap - 1)], make_uint2(info, __float_as_uint(v)));
            };
            const uint32_t eskip = (grp ^ it) & 1u;  // first tile of this item that belongs to this group
            for (uint32_t t = t0 + eskip; t < t1; t += 2) {
                const uint32_t tl = t - t0, itt = it + tl;
                // s = accumulator stage; sb = index of the hand-over barriers of this tile (TSA: per pipeline, four deep)
                const uint32_t s = TSA ? itt % 3u : itt & (kTcAccStages - 1);
                const uint32_t sb = TSA ? grp * 4u + ((itt >> 1) & 3u) : s;
                const uint32_t ph = TSA ? (itt >> 3) & 1 : (itt / kTcAccStages) & 1;
                const float Pnew = lds_volatile_f(&s_P[row]);  // in flight during the wait
                // NB: the smallest norm term of each of the tile's four groups (uniform load, in flight during the wait too)
                float gsc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (NB) {
                    const float4 gm = __ldg(reinterpret_cast<const float4*>(p.gmin + g_list + t * kTcTileGroups));
                    gsc[0] = gm.x * tAones; gsc[1] = gm.y * tAones; gsc[2] = gm.z * tAones; gsc[3] = gm.w * tAones;
                }
                { TC_T0(); mbar_wait(&bar_tfull[sb], ph); if (warp == 1) TC_ACC(8); }
                tc_fence_after();
                const uint32_t ng = min((uint32_t)kTcTileGroups, ngl - t * kTcTileGroups);
                const bool active = valid;
                P = fminf(P, Pnew);
                if (p.flags & 2u) {  // ablation (timing only, wrong answers): the epilogue drains nothing
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_tempty[sb]), 0u));
                        else mbar_arrive(&bar_tempty[sb]);
                    }
                    continue;
                }
                // The tile's 128 columns come out of tensor memory in four loads of 32, double-buffered in registers: while one
                // chunk goes through the min tree the next load is in flight, so only the first load's latency is exposed
                // (tcgen05.wait::ld waits for ALL outstanding loads, hence one load ahead, not more).  Once the fourth load has
                // landed the accumulator stage is handed back to the MMA warp -- before the last chunk is looked at.
                uint32_t ra[32], rb[32];
                const uint32_t tb = tmem_base + ((uint32_t)(quarter * 32) << 16) + s * 128;
                auto chunk = [&](const uint32_t cb, const uint32_t (&r)[32]) -> float {
                    float tv[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) tv[j] = __uint_as_float(r[j]);
                    const float m = min32(tv);
                    if (!NB && p.mode == 2) return cb < ng ? m : kInf;  // bounds pass: the minimum of the 32-column group (+inf past the list)
                    // (NB: m + gmin <= every accumulator of the group + its own norm term, and rounding is monotonic)
                    const float gmn = NB ? gsc[cb & 3u] : 0.0f;
#ifdef VIDX_TC_ABLATE
                    if (p.flags & 0x2000u) return r[0] == 0x12345678u ? 1.0f : 0.0f;  // timing only: loads, no min tree, no hits
                    if ((p.flags & 0x1000u) && m > -1e38f) return m;                    // timing only: min tree, never the hit path
#endif
                    if (active && cb < ng && (NB ? m + gmn : m) <= P) {
                        // rare path: this row has columns inside its bound (as of the latest bound)
#ifdef VIDX_TC_TIMING
                        const long long _th0 = clock64();
#endif
                        P = fminf(P, lds_volatile_f(&s_P[row]));
                        uint32_t mask = 0;
#pragma unroll
                        for (int j = 0; j < 32; j++) mask |= ((NB ? tv[j] + gmn : tv[j]) <= P) ? (1u << j) : 0u;
                        while (mask) {
                            const int j = __ffs(mask) - 1;
                            mask &= mask - 1;
                            // tv[j] for a run-time j: a 5-level multiplexer tree (31 selects, depth 5) instead of a chain of 31
                            float s16[16], s8[8], s4[4];
#pragma unroll
                            for (int i = 0; i < 16; i++) s16[i] = (j & 16) ? tv[16 + i] : tv[i];
#pragma unroll
                            for (int i = 0; i < 8; i++) s8[i] = (j & 8) ? s16[8 + i] : s16[i];
#pragma unroll
                            for (int i = 0; i < 4; i++) s4[i] = (j & 4) ? s8[4 + i] : s8[i];
                            const float s2a = (j & 2) ? s4[2] : s4[0], s2b = (j & 2) ? s4[3] : s4[1];
                            float v = (j & 1) ? s2b : s2a;
                            // NB: the column's own norm term makes it the filter value the ninth MMA would have produced
                            if (NB) v += __ldg(&p.vnorm32[(size_t)(g_list + t * kTcTileGroups + cb) * 32u + (uint32_t)j]) * tAones;
                            if (v <= P) {  // the bound may have shrunk since the mask was built
                                push(kEntValid | (tl << 14) | ((cb * 32u + (uint32_t)j) << 7) | (uint32_t)row, v);
                                // flood control (cold or very loose bound): the k-th smallest value this thread queued
                                // in this item bounds the row's k-th best at once, without the selector's latency.  A warm row
                                // queues a value or two per item and needs none of it: the insertion (a dependent chain of KR
                                // min / max pairs that the other 31 lanes of the warp wait for) starts with the third value --
                                // the k smallest of a subset of the queued values still bound the row's k-th best.
                                if (++nhit > kTcFloodAfter && v < lr[0]) {
                                    lr[0] = v;
#pragma unroll
                                    for (int i = 0; i + 1 < KR; i++) {
                                        const float hi_v = fmaxf(lr[i], lr[i + 1]), lo_v = fminf(lr[i], lr[i + 1]);
                                        lr[i] = hi_v;
                                        lr[i + 1] = lo_v;
                                    }
                                    P = fminf(P, lr[0] + delta);
                                }
                            }
                        }
#ifdef VIDX_TC_TIMING
                        if (p.dbg) atomicAdd(&p.dbg[16 * blockIdx.x + 15], (unsigned long long)(clock64() - _th0));  // thread-cycles in the rare path
#endif
                    }
                    return m;
                };
                tc_ld32_issue(tb, ra);
                { TC_T0(); tc_ld_wait(ra); if (warp == 1) TC_ACC(14); }
                tc_ld32_issue(tb + 32, rb);
                const float m0 = chunk(0, ra);
                { TC_T0(); tc_ld_wait(rb); if (warp == 1) TC_ACC(14); }
                tc_ld32_issue(tb + 64, ra);
                const float m1 = chunk(1, rb);
                { TC_T0(); tc_ld_wait(ra); if (warp == 1) TC_ACC(14); }
                tc_ld32_issue(tb + 96, rb);
                const float m2 = chunk(2, ra);
                { TC_T0(); tc_ld_wait(rb); if (warp == 1) TC_ACC(14); }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_tempty[sb]), 0u));  // the issuer lives in the leader CTA
                    else mbar_arrive(&bar_tempty[sb]);
                }
                const float m3 = chunk(3, rb);
                if (p.mode == 2 && valid) *reinterpret_cast<float4*>(p.submin + (size_t)(submin_row0 + t) * 4) = make_float4(m0, m1, m2, m3);
            }
            it += t1 - t0;
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                atomicAdd(&s_misc[4], 1u);
            }
            asm volatile("bar.sync 1, %0;" ::"n"((kTcEpiWarps + kTcSelectors) * 32) : "memory");
            // merge this item's k smallest into the shared set (distinct values only: a value both sides
            // already hold must not be counted twice; dropping a legitimately equal value only loosens the bound)
            if (p.mode != 2 && valid && grp == 0 && s_impr[row] && !local_sets) {
              const uint32_t q = qi.x;
              const float base_t = s_base[row];
              float r[KR];
#pragma unroll
              for (int i = 0; i < KR; i++) r[i] = s_r[row * kRS + i];
              // insert v into the descending array g (drops the current largest)
              auto insert_desc = [&](float (&g)[KR], float v) {
                  g[0] = v;
#pragma unroll
                  for (int i = 0; i + 1 < KR; i++) {
                      float hi_v = fmaxf(g[i], g[i + 1]), lo_v = fminf(g[i], g[i + 1]);
                      g[i] = hi_v;
                      g[i + 1] = lo_v;
                  }
              };
              // The row's own set (what it read at set-up plus this item's finds) already bounds its k-th best: publish that
              // without the lock.  The union with the other CTAs' sets is tighter still, but it is only worth a short wait:
              // with few query tiles a dozen CTAs finish items of the SAME rows at the same moment, and queueing on the rows'
              // locks (and the set-up's seqlock reads behind them) cost more than the survivors a merge saves -- a row
              // whose lock stays taken skips the merge; its values are not lost to the bound, only to the union.
              if (r[0] < kInf) {
                  float U = fmaxf((r[0] + delta) * tInvS + base_t, 0.0f);
                  U = U + 1e-5f * U;
                  atomicMin(&p.gthr_bits[q], __float_as_uint(U));
              }
              const int merge_tries = ((p.flags >> 4) & 0xffu) ? (int)((p.flags >> 4) & 0xffu) : kTcMergeTries;  // (flags bits 4-11: A/B override)
              bool done = false;
#ifdef VIDX_TC_ABLATE
              if (p.flags & 0x10000u) done = true;  // (valid answers) no union with the other CTAs' sets at item end
#endif
              for (int tries = 0; !done && tries < merge_tries; tries++) {
                // the critical section runs INSIDE the retry loop: a lane that holds a lock always finishes
                // and releases it before it waits for the other lanes of its warp
                if (atomicCAS(&p.glock[q], 0u, 1u) != 0u) {
                    __nanosleep(64);
                    continue;
                }
                __threadfence();
                float g[KR];
#pragma unroll
                for (int i = 0; i < KR; i++) g[i] = i < (int)p.k ? __ldcg(&p.gtop[(size_t)q * p.k + i]) : -kInf;
                bool changed = false;
#pragma unroll
                for (int i = KR - 1; i >= 0; i--) {
                    const float v = r[i];
                    if (v > -kInf && v < g[0]) {
                        bool dup = false;
#pragma unroll
                        for (int j = 0; j < KR; j++) dup = dup || (g[j] == v);
                        if (!dup) {
                            insert_desc(g, v);
                            changed = true;
                        }
                    }
                }
                if (changed) {
                    atomicAdd(p.gver + q, 1u);  // odd: update in progress
                    __threadfence();
#pragma unroll
                    for (int i = 0; i < KR; i++)
                        if (i < (int)p.k) __stcg(&p.gtop[(size_t)q * p.k + i], g[i]);
                    __threadfence();
                    atomicAdd(p.gver + q, 1u);  // even: consistent again
                    if (g[0] < kInf) {
                        float U = fmaxf((g[0] + delta) * tInvS + base_t, 0.0f);
                        U = U + 1e-5f * U;
                        atomicMin(&p.gthr_bits[q], __float_as_uint(U));
                    }
                }
                __threadfence();
                atomicExch(&p.glock[q], 0u);
                done = true;
              }
            }
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();  // neither CTA frees tensor memory (or exits) while the other may still use the pair's
    else __syncthreads();
#ifdef VIDX_TC_TIMING
    if (tid == 0 && p.dbg) p.dbg[16 * blockIdx.x + 13] += (unsigned long long)(clock64() - _tk0);
#endif
    if (warp == 1) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// bounds pass (few tile visits per query: the regime where the scan is HBM-bound and every row of a work item would
// start cold): a first pass of the filter kernel writes only the minimum of every 32 columns of every (query, probed
// tile).  The k-th smallest of a query's minima is an upper bound of its k-th smallest filter value (they are values of
// k distinct vectors), and a tight one; the main pass then runs with final bounds and only collects survivors.
// ------------------------------------------------------------------------------------------
__global__ void pair_tiles_kernel(const uint32_t* __restrict__ probes, size_t npairs, const uint32_t* __restrict__ list_ngroups,
                                  uint32_t* __restrict__ pair_tiles) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const uint32_t l = probes[i];
    pair_tiles[i] = l == kNoRow ? 0u : (list_ngroups[l] + kTcTileGroups - 1) / kTcTileGroups;
}
__global__ void submin_rows_kernel(const uint32_t* __restrict__ pair_off, uint32_t nprobe, uint32_t nq, uint64_t* __restrict__ row_off,
                                   uint32_t* __restrict__ row_len) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint32_t a = pair_off[(size_t)q * nprobe], b = pair_off[(size_t)(q + 1) * nprobe];
    row_off[q] = (uint64_t)a * 4;
    row_len[q] = (b - a) * 4;
}
// The k smallest minima (accumulator units; values of k distinct vectors) become the query's top-k set, stored descending:
// the main pass reads its bound from the set's largest element.  Queries with fewer than k minima keep the empty set
// (bound +inf).
__global__ void bounds_apply_kernel(const float* __restrict__ sel_val /* [nq][k] ascending */, uint32_t nq, uint32_t k,
                                    float* __restrict__ gtop) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * k) return;
    const float kth = sel_val[(size_t)(i / k) * k + (k - 1)];
    if (kth < __int_as_float(0x7f800000)) gtop[i] = sel_val[(size_t)(i / k) * k + (k - 1 - i % k)];
}

// Multi-GPU: a rank's bounds are those of its own part of the index -- the k-th best of an eighth of the vectors is about the
// (8k)-th best of all of them, so every rank would let ~4x more candidates through its filter than one GPU does.  After the bounds
// launch every rank turns the k smallest minima of each query (sel_val, accumulator units, ascending) into UPPER bounds of the
// exact distances of k distinct vectors (real units: the scales differ between ranks), the ranks all-gather them, and the k-th
// smallest of the union -- an upper bound of the global k-th best distance -- becomes the query's published bound on every rank.
__global__ void bounds_to_ub_kernel(const float* __restrict__ sel_val, uint32_t nq, uint32_t k, const float* __restrict__ qnorm,
                                    const TcScale* __restrict__ scale, float vn_max, float* __restrict__ ub) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * k) return;
    const uint32_t q = i / k;
    const float v = sel_val[i], kInf = __int_as_float(0x7f800000);
    float u = kInf;
    if (scale->ok && v < kInf) {
        const float qn = qnorm[q], cabs = scale->c_abs;
        const float base_t = (1.0f - kTcEps) * qn - cabs;
        const float delta = 2.0f * kTcEps * (qn + vn_max) + 2.0f * cabs;  // real units (scan_tc_kernel keeps it times S)
        u = fmaxf((v * scale->invS + delta) + base_t, 0.0f);
        u = u + 1e-5f * u;
    }
    ub[i] = u;
}
__global__ void bounds_merge_kernel(const float* __restrict__ ub_all /* [world][nq][k] */, uint32_t world, uint32_t nq, uint32_t k,
                                    uint32_t* __restrict__ gthr_bits) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float kInf = __int_as_float(0x7f800000);
    float best[32];  // the k smallest so far, ascending (k <= 32)
#pragma unroll
    for (int i = 0; i < 32; i++) best[i] = kInf;
    for (uint32_t r = 0; r < world; r++) {
        const float* row = ub_all + ((size_t)r * nq + q) * k;
        for (uint32_t i = 0; i < k; i++) {
            float v = row[i];
            if (!(v < best[k - 1])) break;  // rows ascend: nothing smaller follows
#pragma unroll
            for (int j = 0; j < 32; j++) {  // insertion: v bubbles up from its place
                if ((uint32_t)j < k && v < best[j]) {
                    const float t = best[j];
                    best[j] = v;
                    v = t;
                }
            }
        }
    }
    if (best[k - 1] < kInf) atomicMin(&gthr_bits[q], __float_as_uint(best[k - 1]));
}

// ------------------------------------------------------------------------------------------
// finalize: exact distances of the survivors + merge with the exact-path slots -> top-k
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool lex_less(float d, unsigned long long key, float d2, unsigned long long key2) {
    return d < d2 || (d == d2 && key < key2);
}
__device__ __forceinline__ void warp_insert_lex64(float cd, unsigned long long ck, float& my_d, unsigned long long& my_k, int lane) {
    unsigned m = __ballot_sync(kFull, lex_less(my_d, my_k, cd, ck));
    int pos = __popc(m);
    float up_d = __shfl_up_sync(kFull, my_d, 1);
    unsigned long long up_k = __shfl_up_sync(kFull, my_k, 1);
    if (lane == pos) { my_d = cd; my_k = ck; }
    else if (lane > pos) { my_d = up_d; my_k = up_k; }
}
// The reference's exact distance of one stored row (utils.rs:28-30), from the interleaved store.
__device__ __forceinline__ float exact_row_distance(const float4* __restrict__ vecs, int Dq, uint32_t row,
                                                    const float4* __restrict__ q4) {
    const float4* pv = vecs + f4_row_base(row, Dq);
    float a = 0.0f;
#pragma unroll 4
    for (int c = 0; c < Dq; c++) {
        float4 v = __ldg(pv + (size_t)c * kSuper);
        float4 q = __ldg(q4 + c);
        a = sqdiff_acc(a, q.x, v.x);
        a = sqdiff_acc(a, q.y, v.y);
        a = sqdiff_acc(a, q.z, v.z);
        a = sqdiff_acc(a, q.w, v.w);
    }
    return a;
}

// One warp per query, or -- small batches, where a lone warp's chain of gathers is the run time -- the eight warps of a
// block share a query: each takes every eighth batch of survivors / slots and keeps its own top-k, warp 0 merges the eight
// lists.  (distance, key) is a strict total order, so the result does not depend on the split.
constexpr int kFinWarps = 8;
__global__ void __launch_bounds__(kFinWarps * 32) finalize_kernel(FinalizeParams p) {
    const uint32_t wib = threadIdx.x >> 5;  // warp in block
    const uint32_t wpq = p.wpq;             // warps per query: 1 or kFinWarps
    const uint32_t q = wpq == 1 ? blockIdx.x * kFinWarps + wib : blockIdx.x;
    const uint32_t sub = wpq == 1 ? 0u : wib;
    int lane = threadIdx.x & 31;
    if (q >= p.nq) return;
    const uint32_t k = p.k;
    float fd = __int_as_float(0x7f800000);
    unsigned long long fk = ~0ull;
    // (1) tensor-core survivors of this query (skipped when its buffer overflowed: the exact path redid it)
    if (p.cand) {
        uint32_t n = p.cand_cnt[q];
        bool ovf = p.overflow[q] != 0 || n > p.capq;
        // one-list tables (coarse quantization) have no exact kernels behind them: a query the filter handed back
        // (a batch that cannot be scaled into fp16 range) is answered here by checking every row exactly
        const bool brute = ovf && p.brute_rows != 0;
        if (brute) n = p.brute_rows;
        if (!ovf || brute) {
            const float4* q4 = p.xq4 + (size_t)q * p.Dq;
            // Survivors were collected under bounds that kept shrinking: most of the early ones lie above the query's final
            // bound.  Their filter values (lower bounds of the distance) are compared with it first, and the rest is
            // compacted through shared memory so that the distance gathers run on full batches.
            __shared__ unsigned long long s_keep[kFinWarps][64];
            unsigned long long* keep = s_keep[wib];
            const bool filt = !brute && p.cand_val != nullptr;
            float U = __int_as_float(0x7f800000), base_t = 0.0f, invS = 0.0f;
            if (filt) {
                U = __uint_as_float(p.gthr_bits[q]);
                base_t = (1.0f - kTcEps) * p.qnorm[q] - p.scale->c_abs;
                invS = p.scale->invS;
            }
            uint32_t pend = 0;
            const uint32_t step = 32u * wpq;
            for (uint32_t base0 = 32u * sub; base0 < n || pend; base0 += step) {
                unsigned long long key = ~0ull;
                bool have = false;
                if (filt) {
                    const uint32_t i0 = base0 + lane;
                    bool kp = false;
                    unsigned long long kk0 = ~0ull;
                    if (i0 < n) {
                        kk0 = p.cand[(size_t)q * p.capq + i0];
                        kp = __fmaf_rn(p.cand_val[(size_t)q * p.capq + i0], invS, base_t) <= U;
                    }
                    const unsigned km = __ballot_sync(kFull, kp);
                    if (kp) keep[pend + __popc(km & ((1u << lane) - 1u))] = kk0;
                    pend += __popc(km);
                    __syncwarp();
                    if (pend < 32 && base0 + step < n) continue;  // gather more before paying for a batch of distances
                    const uint32_t take = min(pend, 32u);
                    have = (uint32_t)lane < take;
                    if (have) key = keep[lane];
                    __syncwarp();
                    if (pend > 32 && (uint32_t)lane < pend - 32) keep[lane] = keep[32 + lane];  // (pend <= 63)
                    pend -= take;
                    __syncwarp();
                    if (!take) continue;
                } else {
                    const uint32_t i = base0 + lane;
                    have = i < n;
                    key = have ? (brute ? (unsigned long long)i : p.cand[(size_t)q * p.capq + i]) : ~0ull;
                }
                float d = have ? exact_row_distance(p.vecs, p.Dq, (uint32_t)key, q4) : __int_as_float(0x7f800000);
                float td = __shfl_sync(kFull, fd, k - 1);
                unsigned long long tk = __shfl_sync(kFull, fk, k - 1);
                unsigned m = __ballot_sync(kFull, have && lex_less(d, key, td, tk));
                while (m) {
                    int src = __ffs(m) - 1;
                    m &= m - 1;
                    float cd = __shfl_sync(kFull, d, src);
                    unsigned long long ck = __shfl_sync(kFull, key, src);
                    td = __shfl_sync(kFull, fd, k - 1);
                    tk = __shfl_sync(kFull, fk, k - 1);
                    if (lex_less(cd, ck, td, tk)) warp_insert_lex64(cd, ck, fd, fk, lane);
                }
            }
        }
    }
    // (2) slots of the exact scan kernels (sorted by (distance, row); slot order = (probe rank, segment))
    if (p.slot_off) {
        uint32_t s0 = p.slot_off[(size_t)q * p.nprobe], s1 = p.slot_off[(size_t)(q + 1) * p.nprobe];
        for (uint32_t s = s0 + sub; s < s1; s += wpq) {
            float d = __int_as_float(0x7f800000);
            uint32_t r = kNoRow;
            if (lane < (int)k) {
                d = p.slot_d[(size_t)s * k + lane];
                r = p.slot_r[(size_t)s * k + lane];
            }
            unsigned long long key = ((unsigned long long)p.slot_rank[s] << 32) | r;
            float td = __shfl_sync(kFull, fd, k - 1);
            unsigned long long tk = __shfl_sync(kFull, fk, k - 1);
            unsigned m = __ballot_sync(kFull, r != kNoRow && lex_less(d, key, td, tk));
            while (m) {
                int src = __ffs(m) - 1;
                m &= m - 1;
                float cd = __shfl_sync(kFull, d, src);
                unsigned long long ck = __shfl_sync(kFull, key, src);
                td = __shfl_sync(kFull, fd, k - 1);
                tk = __shfl_sync(kFull, fk, k - 1);
                if (lex_less(cd, ck, td, tk)) warp_insert_lex64(cd, ck, fd, fk, lane);
            }
        }
    }
    if (wpq > 1) {
        __shared__ float s_md[kFinWarps][32];
        __shared__ unsigned long long s_mk[kFinWarps][32];
        s_md[wib][lane] = fd;
        s_mk[wib][lane] = fk;
        __syncthreads();  // (all warps of the block belong to query q: nobody returned early)
        if (wib != 0) return;
        for (uint32_t w = 1; w < wpq; w++) {
            const float d = s_md[w][lane];
            const unsigned long long key = s_mk[w][lane];
            float td = __shfl_sync(kFull, fd, k - 1);
            unsigned long long tk = __shfl_sync(kFull, fk, k - 1);
            unsigned m = __ballot_sync(kFull, lane < (int)k && key != ~0ull && lex_less(d, key, td, tk));
            while (m) {
                int src = __ffs(m) - 1;
                m &= m - 1;
                float cd = __shfl_sync(kFull, d, src);
                unsigned long long ck = __shfl_sync(kFull, key, src);
                td = __shfl_sync(kFull, fd, k - 1);
                tk = __shfl_sync(kFull, fk, k - 1);
                if (lex_less(cd, ck, td, tk)) warp_insert_lex64(cd, ck, fd, fk, lane);
            }
        }
    }
    if (lane < (int)k) {
        size_t o = (size_t)q * p.kout + lane;
        bool ok = fk != ~0ull;
        uint32_t row = (uint32_t)fk;
        p.D[o] = ok ? fd : __int_as_float(0x7f800000);
        if (p.I) p.I[o] = ok ? (p.row_ext ? (int64_t)p.row_ext[row] : (int64_t)row) : -1;
        if (p.out_rows) p.out_rows[o] = ok ? row : kNoRow;
        if (p.out_keys) {
            const uint32_t rank = (uint32_t)(fk >> 32);
            p.out_keys[o] = ok ? ((unsigned long long)rank << 32) | (uint32_t)(row + p.list_rowdelta[p.probes[(size_t)q * p.nprobe + rank]])
                               : ~0ull;
        }
    }
}

// ====================================================================================
// launchers
// ====================================================================================
// the query tile (4 columns per chunk) + 8 columns of the norm step fit behind three accumulator stages (D <= 240)
bool tc_tsa_supported(int Dh) { return (uint32_t)Dh * 4u + 8u <= (uint32_t)kTcTmemCols - kTcACol; }
// (D <= 2048: the filter's error budget covers the fp32 accumulation of that many products, see the header comment)
bool tc_supported(int D, uint32_t k) {
    if (k == 0 || k > 32 || D < 1 || D > 2048) return false;
    return tc_smem_layout(tc_dh(D), k <= 8 ? 8 : (k <= 16 ? 16 : 32)).total <= kTcSmemMax;
}
bool tc_streams_a(int D, uint32_t k) {
    return tc_supported(D, k) && tc_smem_layout(tc_dh(D), k <= 8 ? 8 : (k <= 16 ? 16 : 32)).stream_a != 0;
}
size_t tc_atile_rows_cap(size_t npairs, size_t nlist) {
    // sum over the probed lists of their query count rounded up to 8
    return std::min(8 * npairs, npairs + 7 * std::min(nlist, npairs)) + 8;
}
void launch_tc_atiles(const uint32_t* list_cnt, const uint32_t* list_qoff, const uint2* list_qlist, uint32_t nlist, size_t npairs_cap,
                      const float4* xq4, int Dq, int Dh, const TcScale* scale, uint32_t* rows8, uint32_t* a_rowoff, uint32_t* scan_tmp,
                      uint4* a_tiles, cudaStream_t st) {
    if (!nlist || !npairs_cap) return;
    tc_arows_kernel<<<(unsigned)ceil_div(nlist, 256), 256, 0, st>>>(list_cnt, nlist, rows8);
    VIDX_LAUNCHED();
    exclusive_scan_u32(rows8, a_rowoff, nlist, scan_tmp, st);
    tc_atile_kernel<<<(unsigned)ceil_div(npairs_cap, 32), 256, 0, st>>>(list_cnt, list_qoff, list_qlist, a_rowoff, nlist, xq4, Dq, Dh, scale,
                                                                        a_tiles);
    VIDX_LAUNCHED();
}
void launch_row_norms(const float4* vecs, int Dq, const uint32_t* row_src, size_t nrows, float* vn_true, uint32_t* stats,
                      cudaStream_t st) {
    if (!nrows) return;
    row_norm_kernel<<<(unsigned)ceil_div(nrows, 256), 256, 0, st>>>(vecs, Dq, row_src, nrows, vn_true, stats);
    VIDX_LAUNCHED();
}
void launch_convert16(const float4* vecs, int Dq, int Dh, const uint32_t* row_src, size_t nrows, const float* vn_true, int sv, int g,
                      uint4* vecs16, uint4* vnorm, cudaStream_t st, const uint32_t* perm, float* vnorm32, float* gmin) {
    if (!nrows) return;
    convert16_kernel<<<(unsigned)ceil_div(nrows * (size_t)Dh, 256), 256, 0, st>>>(vecs, Dq, Dh, row_src, nrows, vn_true, ldexpf(1.0f, sv),
                                                                                ldexpf(1.0f, 2 * sv - g), perm, vecs16, vnorm, vnorm32);
    VIDX_LAUNCHED();
    if (vnorm32 && gmin) {
        group_min_norm_kernel<<<(unsigned)ceil_div(nrows, 256), 256, 0, st>>>(vnorm32, nrows / 32, gmin);
        VIDX_LAUNCHED();
    }
}
void launch_query_norms(const float4* xq4, int Dq, uint32_t nq, uint32_t k, float* qn, uint32_t* gthr_bits, uint32_t* cand_cnt,
                        uint32_t* overflow, float* gtop, uint32_t* glock, uint32_t* stats, cudaStream_t st) {
    if (!nq) return;
    query_norm_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, st>>>(xq4, Dq, nq, k, qn, gthr_bits, cand_cnt, overflow, gtop, glock, stats);
    VIDX_LAUNCHED();
}
void launch_tc_scale(const uint32_t* qstats, int sv, int g, int D, float vmax, float vn_max, TcScale* out, cudaStream_t st) {
    tc_scale_kernel<<<1, 1, 0, st>>>(qstats, sv, g, D, vmax, vn_max, out);
    VIDX_LAUNCHED();
}
void launch_tc_count(const uint32_t* probes, size_t npairs, uint32_t nprobe, uint32_t max_rank, const uint2* list_seg,
                     uint32_t* list_cnt, cudaStream_t st) {
    if (!npairs) return;
    tc_count_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, st>>>(probes, npairs, nprobe, max_rank, list_seg, list_cnt);
    VIDX_LAUNCHED();
}
void launch_tc_fill(const uint32_t* probes, size_t npairs, uint32_t nprobe, uint32_t max_rank, const uint2* list_seg,
                    const uint32_t* list_qoff, uint32_t* list_cur, uint2* list_qlist, cudaStream_t st) {
    if (!npairs) return;
    tc_fill_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, st>>>(probes, npairs, nprobe, max_rank, list_seg, list_qoff,
                                                                     list_cur, list_qlist);
    VIDX_LAUNCHED();
}
static int tc_num_sms() { return device_num_sms(); }
// tuning knobs for A/B runs on the GPU box (environment, read per launch; unset = the defaults the bench line is quoted on)
static long tc_env(const char* name, long dflt) {
    const char* v = getenv(name);
    return v && *v ? strtol(v, nullptr, 0) : dflt;
}
void launch_tc_items(const uint32_t* list_cnt, const uint32_t* list_ngroups, uint32_t nlist, unsigned long long* total,
                     uint32_t seed_tiles, uint32_t* chunk_out, uint32_t* items_per_list, bool pair, cudaStream_t st) {
    if (!nlist) return;
    const uint32_t qrows = pair ? 2u * kTcM : (uint32_t)kTcM;  // rows of a work item of the main pass
    if (!seed_tiles) {
        tc_work_kernel<<<(unsigned)ceil_div(nlist, 256), 256, 0, st>>>(list_cnt, list_ngroups, nlist, qrows, total);
        VIDX_LAUNCHED();
    }
    // work-item sizing: an item costs ~6 us of set-up and drain beyond its tiles (0.55 us each), so a batch is cut into
    // about items_per_sm items per SM but never into chunks of fewer than min_chunk tiles
    const uint32_t items_per_sm = (uint32_t)tc_env("VIDX_ITEMS_PER_SM", VIDX_ITEMS_PER_SM);
    const uint32_t min_chunk = (uint32_t)tc_env("VIDX_MIN_CHUNK_TILES", VIDX_MIN_CHUNK_TILES);
    tc_items_kernel<<<(unsigned)ceil_div(nlist, 256), 256, 0, st>>>(list_cnt, list_ngroups, nlist, total,
                                                                   (uint32_t)(pair ? tc_num_sms() / 2 : tc_num_sms()), seed_tiles,
                                                                   std::max(1u, items_per_sm), std::max(1u, min_chunk), qrows, chunk_out,
                                                                   items_per_list);
    VIDX_LAUNCHED();
}
void launch_tc_expand(const uint32_t* list_cnt, const uint32_t* list_ngroups, const uint32_t* list_g0, const uint32_t* list_qoff,
                      const uint32_t* item_off, const uint32_t* chunk_tiles, uint32_t nlist, uint32_t seed_tiles, TcItem* items,
                      bool pair, const uint32_t* a_rowoff, cudaStream_t st) {
    if (!nlist) return;
    tc_expand_kernel<<<nlist, 64, 0, st>>>(list_cnt, list_ngroups, list_g0, list_qoff, item_off, chunk_tiles, nlist, seed_tiles,
                                           pair ? 2u * kTcM : (uint32_t)kTcM, a_rowoff, items);
    VIDX_LAUNCHED();
}
template <int KR, bool PAIR, bool TSA = false, bool SA = false, bool NB = false>
static void launch_scan_tc_kr(const TcParams& p, size_t smem, cudaStream_t st) {
    static PerDeviceSize attr;  // the opt-in is per device
    if (attr.needs(smem)) {
        VIDX_CUDA(cudaFuncSetAttribute(scan_tc_kernel<KR, PAIR, TSA, SA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr.set(smem);
    }
    if (PAIR) {
        // clusters of two CTAs: the pair lands on the two SMs of one TPC, which is what tcgen05 cta_group::2 needs
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(tc_num_sms() / 2) * 2);
        cfg.blockDim = dim3(kTcThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        VIDX_CUDA(cudaLaunchKernelEx(&cfg, scan_tc_kernel<KR, PAIR, TSA, SA, NB>, p));
    } else {
        scan_tc_kernel<KR, PAIR, TSA, SA, NB><<<tc_num_sms(), kTcThreads, smem, st>>>(p);
    }
    VIDX_LAUNCHED();
}
void launch_scan_tc(const TcParams& p, cudaStream_t st) {
    const bool pair = p.pair != 0;
    const int kr = p.k <= 8 ? 8 : (p.k <= 16 ? 16 : 32);
    const TcSmemLayout L = tc_smem_layout(p.Dh, kr, pair);
    const size_t smem = L.total;
    if (L.stream_a) {
        if (!p.a_tiles) throw ApiError(6, "tensor-core scan: streamed query tiles were not prepared");
        if (kr == 8) launch_scan_tc_kr<8, false, false, true>(p, smem, st);
        else if (kr == 16) launch_scan_tc_kr<16, false, false, true>(p, smem, st);
        else launch_scan_tc_kr<32, false, false, true>(p, smem, st);
    } else if (pair) {
        if (kr == 8) launch_scan_tc_kr<8, true>(p, smem, st);
        else if (kr == 16) launch_scan_tc_kr<16, true>(p, smem, st);
        else launch_scan_tc_kr<32, true>(p, smem, st);
    } else if (p.tsa && tc_tsa_supported(p.Dh)) {
        if (kr == 8) launch_scan_tc_kr<8, false, true>(p, smem, st);
        else if (kr == 16) launch_scan_tc_kr<16, false, true>(p, smem, st);
        else launch_scan_tc_kr<32, false, true>(p, smem, st);
    } else if (p.nb && p.mode == 0 && p.gmin && p.vnorm32) {
        if (kr == 8) launch_scan_tc_kr<8, false, false, false, true>(p, smem, st);
        else if (kr == 16) launch_scan_tc_kr<16, false, false, false, true>(p, smem, st);
        else launch_scan_tc_kr<32, false, false, false, true>(p, smem, st);
    } else {
        if (kr == 8) launch_scan_tc_kr<8, false>(p, smem, st);
        else if (kr == 16) launch_scan_tc_kr<16, false>(p, smem, st);
        else launch_scan_tc_kr<32, false>(p, smem, st);
    }
}
// The fp16 shadow store as a 2-D tensor of 8-byte elements: dim 0 = one 2 KB chunk block (256 elements: one 16-byte chunk of
// 128 vectors), dim 1 = chunk blocks (Dh per 128-vector tile, tiles consecutive).  Box = 128 elements x 16 blocks: the 64
// vectors of one CTA of a pair, 16 chunks deep, landing as [chunk][64 vectors][16 B] -- the tcgen05 K-major operand layout.
void make_shadow_tensor_map(CUtensorMap* out, const void* vecs16, uint64_t nrows, int Dh) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        VIDX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) throw ApiError(6, "cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const uint64_t nblocks = std::max<uint64_t>(1, (nrows + kSuper - 1) / kSuper) * (uint64_t)Dh;
    const cuuint64_t gdim[2] = {256, nblocks};
    const cuuint64_t gstride[1] = {(cuuint64_t)kSuper * 16};
    const cuuint32_t box[2] = {128, (cuuint32_t)kTcStageChunks};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(vecs16), gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw ApiError(6, "cuTensorMapEncodeTiled failed for the fp16 shadow store");
}
void launch_pair_tiles(const uint32_t* probes, size_t npairs, const uint32_t* list_ngroups, uint32_t* pair_tiles, cudaStream_t st) {
    if (!npairs) return;
    pair_tiles_kernel<<<(unsigned)ceil_div(npairs, 256), 256, 0, st>>>(probes, npairs, list_ngroups, pair_tiles);
    VIDX_LAUNCHED();
}
void launch_submin_rows(const uint32_t* pair_off, uint32_t nprobe, uint32_t nq, uint64_t* row_off, uint32_t* row_len, cudaStream_t st) {
    if (!nq) return;
    submin_rows_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(pair_off, nprobe, nq, row_off, row_len);
    VIDX_LAUNCHED();
}
void launch_bounds_apply(const float* sel_val, uint32_t nq, uint32_t k, float* gtop, cudaStream_t st) {
    if (!nq) return;
    bounds_apply_kernel<<<(unsigned)ceil_div((size_t)nq * k, 256), 256, 0, st>>>(sel_val, nq, k, gtop);
    VIDX_LAUNCHED();
}
void launch_bounds_to_ub(const float* sel_val, uint32_t nq, uint32_t k, const float* qnorm, const TcScale* scale, float vn_max, float* ub,
                         cudaStream_t st) {
    if (!nq) return;
    bounds_to_ub_kernel<<<(unsigned)ceil_div((size_t)nq * k, 256), 256, 0, st>>>(sel_val, nq, k, qnorm, scale, vn_max, ub);
    VIDX_LAUNCHED();
}
void launch_bounds_merge(const float* ub_all, uint32_t world, uint32_t nq, uint32_t k, uint32_t* gthr_bits, cudaStream_t st) {
    if (!nq) return;
    bounds_merge_kernel<<<(unsigned)ceil_div(nq, 128), 128, 0, st>>>(ub_all, world, nq, k, gthr_bits);
    VIDX_LAUNCHED();
}
void launch_finalize(const FinalizeParams& p, cudaStream_t st) {
    if (!p.nq) return;
    FinalizeParams q = p;
    if (q.wpq != (uint32_t)kFinWarps) q.wpq = 1;
    finalize_kernel<<<q.wpq == 1 ? (unsigned)ceil_div((size_t)q.nq, kFinWarps) : q.nq, kFinWarps * 32, 0, st>>>(q);
    VIDX_LAUNCHED();
}

}  // namespace vidx

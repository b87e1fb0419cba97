// scan_tc.h -- tensor-core list scan (tcgen05 FP16 pre-filter + exact finalize).
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace vidx {

// Power-of-two scales of one search batch, written on the device by tc_scale_kernel: stored vectors are scaled by
// 2^sv when the fp16 shadow store is built (index constant), queries by 2^sq per batch, both to a largest
// component in [2^6, 2^7), so products and sums stay far inside fp16 / fp32 range.
constexpr int kTcTileGroups = 4;  // groups of 32 vectors per tile of the tensor-core scan: 128 vectors (UMMA N)

struct TcScale {
    float S;        // 2^(sq+sv): an accumulator holds S * (filter value in real units)
    float invS;
    float qmul;     // -2 * 2^sq: query components are multiplied by this before the fp16 conversion
    float a_ones;   // 2^(sq-sv+g): A-side partner of the three norm terms b = (1-eps)|v|^2 * 2^(2sv-g)
    float c_abs;    // absolute error of 2*dot from fp16 subnormals, real units (0 for ordinary data)
    uint32_t ok;    // 0: the batch cannot be scaled into fp16 range; every query is handed to the exact kernels
};

struct TcItem {      // one work item: query tile x list chunk (32 bytes)
    uint32_t t0, t1;     // tile range within the list
    uint32_t qbase;      // first entry of the item's rows in list_qlist
    uint32_t nq_tile;    // rows in use (<= 128; <= 256 for the CTA-pair kernel)
    uint32_t g_list;     // first group of the list (this rank's part)
    uint32_t ngl;        // groups of the list (this rank's part)
    uint32_t valid, pad;  // pad: streamed query tiles (D > 512) -- first row of the item's tile in TcParams::a_tiles
};

struct TcParams {
    CUtensorMap tmap;              // CTA-pair kernel only: the shadow store as [chunk blocks][2 KB] (make_shadow_tensor_map)
    const uint4* vecs16;           // fp16 shadow store: [supergroup][Dh chunks][128 vectors][8 halfs], scaled by 2^sv
    const uint4* vnorm;            // per row 8 halfs: the three fp16 terms of (1-eps)|v|^2 * 2^(2sv-g), then zeros; NaN for padding rows
    int Dh;                        // 16-byte chunks (8 halfs) per vector in the shadow store; even
    int Dq;                        // float4 per query row
    const TcScale* scale;
    const float4* xq4;             // queries, row-major, Dq float4 per row
    const float* qnorm;            // per query |q|^2
    const uint32_t* list_g0;       // first group of each list
    const uint32_t* list_ngroups;  // groups of each list
    const uint32_t* list_cnt;      // queries probing each list (this batch)
    const uint32_t* list_qoff;     // CSR offsets into list_qlist
    const uint2* list_qlist;       // (query, probe rank)
    const uint32_t* item_off;      // nlist+1: prefix of work items per list
    const TcItem* items;           // item_off[nlist] records (tc_expand_kernel)
    uint32_t nlist;
    uint32_t nq;
    uint32_t* work_counter;
    uint32_t* gthr_bits;           // per query: upper bound of the exact k-th best distance (float bits)
    float* gtop;                   // per query: the k smallest filter values found so far by any CTA, descending
    uint32_t* glock;               // per query: spin lock serialising writers of gtop
    uint32_t* gver;                // per query: seqlock version of gtop (odd while being written)
    unsigned long long* cand;      // per query capq survivors: (rank << 32 | row)
    uint32_t* cand_cnt;
    uint32_t* overflow;
    uint32_t capq;
    uint32_t k;
    float vn_max;                  // max |v|^2 over the stored rows
    unsigned long long* dbg;       // role timers (builds with -DVIDX_TC_TIMING only), else NULL
    float* submin;                 // bounds pass: the minimum of every 32 columns of every (query, probed tile): 4 floats per tile
    const uint32_t* pair_off;      // bounds pass: first tile of pair (query * nprobe + rank) in submin; null (seeding bounds pass) =
                                   // (query * seed_ranks + rank) * noinsert_tiles
    uint32_t nprobe;
    float* cand_val;               // optional: filter value (accumulator units) of every survivor, parallel to cand
    uint32_t mode;                 // 0 = main pass, 2 = bounds pass (minima only)
    uint32_t seed_ranks, noinsert_tiles;  // seeding bounds pass: the first noinsert_tiles tiles of each query's seed_ranks nearest lists;
                                   // main pass after it: their values are survivors but never enter the row's set
    uint32_t frozen;               // main pass after a bounds pass: the bounds are final, survivors are only collected
    uint32_t tsa;                  // 1 = query tile in tensor memory (tcgen05.mma with A from TMEM, three accumulator stages), D <= 240
    uint32_t pair;                 // 1 = the CTA-pair kernel (cta_group::2): work items of 256 query rows, clusters of two CTAs
    uint32_t flags;                // bit 0: keep the rows' sets CTA-local in the main pass (no cross-CTA merge at item ends);
                                   // bits 1, 2: timing ablations (VIDX_TC_FLAGS, wrong answers): epilogue / MMAs do nothing;
                                   // bit 3: the producer feeds the two tile pipelines independently instead of in tile order
    const uint32_t* perm;          // shadow row -> row of the fp32 store (rows are sorted by norm inside every segment); NULL = identity
    const float* gmin;             // per group of 32 shadow rows: the smallest norm term (fp32, same scale as vnorm); NaN for groups of padding rows
    const float* vnorm32;          // per shadow row: the norm term as one fp32 value, rounded down; NaN for padding rows
    uint32_t nb;                   // 1 = main pass with epilogue-side norms (scan_tc_kernel<.., NB>): eight MMAs per 128-d tile instead of nine
    const uint4* a_tiles;          // D > 512 only (tc_streams_a): every work item's query tile as fp16 in operand layout (launch_tc_atiles)
};

struct FinalizeParams {
    uint32_t nq, nprobe, k, kout;
    int Dq;
    const float4* vecs;
    const float4* xq4;
    // tensor-core survivors (cand == nullptr: none)
    const unsigned long long* cand;
    const uint32_t* cand_cnt;
    const uint32_t* overflow;
    uint32_t capq;
    // optional: the survivors' filter values (accumulator units) and what converts them -- a survivor whose lower bound
    // exceeds the query's final published bound (gthr) is skipped without a distance computation
    const float* cand_val;
    const uint32_t* gthr_bits;
    const float* qnorm;
    const TcScale* scale;
    uint32_t wpq;         // warps per query: 8 = a block per query (small batches), anything else = one warp per query
    uint32_t brute_rows;  // != 0: a query flagged in `overflow` is answered by checking rows 0..brute_rows-1 exactly (one-list tables)
    // exact-path slots (slot_off == nullptr: none)
    const uint32_t* slot_off;
    const float* slot_d;
    const uint32_t* slot_r;
    const uint32_t* slot_rank;
    const uint64_t* row_ext;
    float* D;
    int64_t* I;
    uint32_t* out_rows;
    // optional (multi-GPU merge): per result (probe rank << 32 | global row); needs the probe lists and, per list,
    // (global row - local row) of this rank's resident part
    unsigned long long* out_keys;
    const uint32_t* probes;
    const uint32_t* list_rowdelta;
};

bool tc_supported(int D, uint32_t k);  // D = vector dimension
bool tc_tsa_supported(int Dh);
// D > 512: the query tile does not fit in shared memory next to the ring and is streamed through it (TcParams::a_tiles)
bool tc_streams_a(int D, uint32_t k);
size_t tc_atile_rows_cap(size_t npairs, size_t nlist);  // rows (of Dh 16-byte chunks) a_tiles needs for a grouping of npairs (query, list) pairs
// rows8, a_rowoff: nlist + 1 entries each; a_tiles: tc_atile_rows_cap(...) * Dh * 16 bytes.  Needs the TcScale of the batch.
void launch_tc_atiles(const uint32_t* list_cnt, const uint32_t* list_qoff, const uint2* list_qlist, uint32_t nlist, size_t npairs_cap,
                      const float4* xq4, int Dq, int Dh, const TcScale* scale, uint32_t* rows8, uint32_t* a_rowoff, uint32_t* scan_tmp,
                      uint4* a_tiles, cudaStream_t st);
int tc_dh(int D);  // chunks per vector of the shadow store (dimension padded to a multiple of 16)
// |v|^2 per row (0 for padding rows) and, in stats[0], the float bits of the largest |component|.
void launch_row_norms(const float4* vecs, int Dq, const uint32_t* row_src, size_t nrows, float* vn_true, uint32_t* stats,
                      cudaStream_t st);
// The fp16 shadow store and the norm terms, given the index scale sv and the norm shift g.
// perm (optional): shadow row -> store row; vnorm32 / gmin (optional): fp32 norm term per shadow row / its minimum per group of 32.
void launch_convert16(const float4* vecs, int Dq, int Dh, const uint32_t* row_src, size_t nrows, const float* vn_true, int sv, int g,
                      uint4* vecs16, uint4* vnorm, cudaStream_t st, const uint32_t* perm = nullptr, float* vnorm32 = nullptr,
                      float* gmin = nullptr);
// stats[0..1]: float bits of max |q component| and max |q|^2 over the batch (zeroed by the caller).
void launch_query_norms(const float4* xq4, int Dq, uint32_t nq, uint32_t k, float* qn, uint32_t* gthr_bits, uint32_t* cand_cnt,
                        uint32_t* overflow, float* gtop, uint32_t* glock, uint32_t* stats, cudaStream_t st);
void launch_tc_scale(const uint32_t* qstats, int sv, int g, int D, float vmax, float vn_max, TcScale* out, cudaStream_t st);
void launch_tc_count(const uint32_t* probes, size_t npairs, uint32_t nprobe, uint32_t max_rank, const uint2* list_seg,
                     uint32_t* list_cnt, cudaStream_t st);
void launch_tc_fill(const uint32_t* probes, size_t npairs, uint32_t nprobe, uint32_t max_rank, const uint2* list_seg,
                    const uint32_t* list_qoff, uint32_t* list_cur, uint2* list_qlist, cudaStream_t st);
void launch_tc_items(const uint32_t* list_cnt, const uint32_t* list_ngroups, uint32_t nlist, unsigned long long* total,
                     uint32_t seed_tiles, uint32_t* chunk_out, uint32_t* items_per_list, bool pair, cudaStream_t st);
void launch_tc_expand(const uint32_t* list_cnt, const uint32_t* list_ngroups, const uint32_t* list_g0, const uint32_t* list_qoff,
                      const uint32_t* item_off, const uint32_t* chunk_tiles, uint32_t nlist, uint32_t seed_tiles, TcItem* items,
                      bool pair, const uint32_t* a_rowoff, cudaStream_t st);
void launch_scan_tc(const TcParams& p, cudaStream_t st);
void make_shadow_tensor_map(CUtensorMap* out, const void* vecs16, uint64_t nrows, int Dh);
void launch_pair_tiles(const uint32_t* probes, size_t npairs, const uint32_t* list_ngroups, uint32_t* pair_tiles, cudaStream_t st);
void launch_submin_rows(const uint32_t* pair_off, uint32_t nprobe, uint32_t nq, uint64_t* row_off, uint32_t* row_len, cudaStream_t st);
void launch_bounds_apply(const float* sel_val, uint32_t nq, uint32_t k, float* gtop, cudaStream_t st);
// multi-GPU bound exchange (see bounds_to_ub_kernel): local k smallest minima -> upper bounds in real units; k-th smallest of all ranks' -> gthr
void launch_bounds_to_ub(const float* sel_val, uint32_t nq, uint32_t k, const float* qnorm, const TcScale* scale, float vn_max, float* ub,
                         cudaStream_t st);
void launch_bounds_merge(const float* ub_all, uint32_t world, uint32_t nq, uint32_t k, uint32_t* gthr_bits, cudaStream_t st);
void launch_finalize(const FinalizeParams& p, cudaStream_t st);

}  // namespace vidx

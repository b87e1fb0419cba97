// chacha_blocks.cpp -- the ChaCha12 keystream of csrc/rng.hpp, eight 64-byte blocks per call with the blocks as SIMD lanes.
// The build's random stream is serial by the reference's semantics (every k-means++ pick and every batch shuffle draws from
// one StdRng, src/kmeans.rs:31,240,722-726), so the generator's speed is on the critical path of a build: configs[2] draws
// ~28 M words.  Plain host code (no CUDA): compiled by the host compiler, AVX2 where the CPU has it (resolved at load time).
#include <cstdint>

namespace vidx {

namespace {
constexpr int kLanes = 8;
typedef uint32_t Row __attribute__((vector_size(32)));   // one state word of eight blocks
template <int R>
inline __attribute__((always_inline)) Row rotl(Row x) { return (x << R) | (x >> (32 - R)); }
inline __attribute__((always_inline)) void quarter(Row& a, Row& b, Row& c, Row& d) {
    a += b; d = rotl<16>(d ^ a);
    c += d; b = rotl<12>(b ^ c);
    a += b; d = rotl<8>(d ^ a);
    c += d; b = rotl<7>(b ^ c);
}
inline __attribute__((always_inline)) Row splat(uint32_t v) { return Row{v, v, v, v, v, v, v, v}; }
}  // namespace

// out[blk * 16 + w] = word w of block (block0 + blk), blk < 8: the order a one-block-at-a-time generator would produce
__attribute__((target_clones("avx2", "default")))
void chacha12_blocks8(const uint32_t key[8], uint64_t block0, uint32_t* out) {
    Row init[16], s[16];
    init[0] = splat(0x61707865u); init[1] = splat(0x3320646eu); init[2] = splat(0x79622d32u); init[3] = splat(0x6b206574u);
    for (int w = 0; w < 8; w++) init[4 + w] = splat(key[w]);
    for (int l = 0; l < kLanes; l++) {
        const uint64_t b = block0 + (uint64_t)l;
        init[12][l] = (uint32_t)b;
        init[13][l] = (uint32_t)(b >> 32);
    }
    init[14] = splat(0u);   // stream id 0 (seed_from_u64)
    init[15] = splat(0u);
    for (int w = 0; w < 16; w++) s[w] = init[w];
    for (int dr = 0; dr < 6; dr++) {  // 12 rounds = 6 (column, diagonal) double rounds
        quarter(s[0], s[4], s[8], s[12]); quarter(s[1], s[5], s[9], s[13]);
        quarter(s[2], s[6], s[10], s[14]); quarter(s[3], s[7], s[11], s[15]);
        quarter(s[0], s[5], s[10], s[15]); quarter(s[1], s[6], s[11], s[12]);
        quarter(s[2], s[7], s[8], s[13]); quarter(s[3], s[4], s[9], s[14]);
    }
    for (int w = 0; w < 16; w++) {
        const Row r = s[w] + init[w];
        for (int l = 0; l < kLanes; l++) out[l * 16 + w] = r[l];
    }
}

}  // namespace vidx

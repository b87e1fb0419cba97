// rng.hpp -- host-side random stream of the product: what the reference draws from
// rand 0.8.5's StdRng (ChaCha12, rand_chacha 0.3.1) seeded with seed_from_u64
// (rand_core 0.6.4).  Every random decision of k-means (k-means++ picks, batch shuffles,
// reseeds, hierarchy seeds; src/kmeans.rs:31,80,170,240,591) is taken on the host with
// this stream, so a given seed yields the same index as the reference's algorithm.
//
// This is an independent implementation: the test oracle under oracle/ carries its own.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

namespace vidx {

class ChaCha12Rng {
  public:
    // SeedableRng::seed_from_u64: the 32-byte key is eight PCG32 outputs.
    explicit ChaCha12Rng(uint64_t seed) {
        uint64_t s = seed;
        for (auto& word : key_) {
            s = s * kPcgMul + kPcgInc;
            uint32_t xs = static_cast<uint32_t>(((s >> 18) ^ s) >> 27);
            unsigned rot = static_cast<unsigned>(s >> 59);
            word = rotr(xs, rot);
        }
    }

    uint32_t next_u32() {
        if (pos_ >= kBufWords) { refill(); pos_ = 0; }
        return buf_[pos_++];
    }

    // BlockRng::next_u64: two consecutive words, low word first, with the straddling
    // case when exactly one word is left in the buffer.
    uint64_t next_u64() {
        if (pos_ + 1 < kBufWords) {
            uint64_t lo = buf_[pos_], hi = buf_[pos_ + 1];
            pos_ += 2;
            return (hi << 32) | lo;
        }
        if (pos_ >= kBufWords) {
            refill();
            pos_ = 2;
            return (static_cast<uint64_t>(buf_[1]) << 32) | buf_[0];
        }
        uint64_t lo = buf_[kBufWords - 1];
        refill();
        pos_ = 1;
        return (static_cast<uint64_t>(buf_[0]) << 32) | lo;
    }

    // Rng::gen_range(0..bound) for usize (64-bit): Lemire-style widening multiply with
    // rand's conservative rejection zone.
    uint64_t below_u64(uint64_t bound) {
        if (bound == 0) return next_u64();
        const uint64_t zone = (bound << __builtin_clzll(bound)) - 1;
        while (true) {
            unsigned __int128 wide = static_cast<unsigned __int128>(next_u64()) * bound;
            if (static_cast<uint64_t>(wide) <= zone) return static_cast<uint64_t>(wide >> 64);
        }
    }
    uint32_t below_u32(uint32_t bound) {
        if (bound == 0) return next_u32();
        const uint32_t zone = (bound << __builtin_clz(bound)) - 1;
        while (true) {
            uint64_t wide = static_cast<uint64_t>(next_u32()) * bound;
            if (static_cast<uint32_t>(wide) <= zone) return static_cast<uint32_t>(wide >> 32);
        }
    }
    // rand::seq index sampling: 32-bit draws whenever the bound fits.
    uint64_t index_below(uint64_t bound) {
        return bound <= 0xffffffffull ? below_u32(static_cast<uint32_t>(bound)) : below_u64(bound);
    }

    // 23 mantissa bits -> [0, 1)
    float unit_float() {
        uint32_t bits = 0x3f800000u | (next_u32() >> 9);
        float f;
        std::memcpy(&f, &bits, sizeof f);
        return f - 1.0f;
    }

    // SliceRandom::shuffle
    template <class T>
    void shuffle(T* a, size_t n) {
        for (size_t i = n; i > 1; --i) {
            size_t j = index_below(i);
            T t = a[i - 1];
            a[i - 1] = a[j];
            a[j] = t;
        }
    }

    // IteratorRandom::choose_multiple over 0..n
    std::vector<uint32_t> choose_multiple(uint32_t n, uint32_t amount) {
        std::vector<uint32_t> r;
        r.reserve(amount);
        uint32_t i = 0;
        while (i < n && r.size() < amount) r.push_back(i++);
        if (r.size() == amount) {
            for (uint64_t seen = amount; i < n; ++i) {
                ++seen;
                uint64_t slot = index_below(seen);
                if (slot < amount) r[slot] = i;
            }
        }
        return r;
    }

    // WeightedIndex::new(weights).sample(): weights non-negative, total > 0.
    // `cum` is scratch (n-1 prefix sums, sequential f32 accumulation).
    size_t weighted_pick(const float* w, size_t n, float total, std::vector<float>& cum) {
        cum.resize(n ? n - 1 : 0);
        float run = w[0];
        for (size_t i = 1; i < n; ++i) {
            cum[i - 1] = run;
            run += w[i];
        }
        // Uniform::new(0, total): shrink scale until the largest sample stays below total
        float scale = total;
        const float max_unit = 0.99999988079071044921875f;  // 1 - 2^-23
        while (scale * max_unit + 0.0f >= total) {
            uint32_t b;
            std::memcpy(&b, &scale, 4);
            --b;
            std::memcpy(&scale, &b, 4);
        }
        const float x = unit_float() * scale + 0.0f;
        // first prefix sum strictly greater than x
        size_t lo = 0, hi = cum.size();
        while (lo < hi) {
            size_t mid = (lo + hi) >> 1;
            if (cum[mid] <= x) lo = mid + 1; else hi = mid;
        }
        return lo;
    }

  private:
    static constexpr uint64_t kPcgMul = 6364136223846793005ull;
    static constexpr uint64_t kPcgInc = 11634580027462260723ull;
    static constexpr int kBufWords = 64;  // four 16-word blocks per refill

    static uint32_t rotr(uint32_t v, unsigned r) { r &= 31; return r ? (v >> r) | (v << (32 - r)) : v; }
    static uint32_t rotl(uint32_t v, unsigned r) { return (v << r) | (v >> (32 - r)); }

    static void quarter(std::array<uint32_t, 16>& s, int a, int b, int c, int d) {
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 16);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 12);
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 8);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 7);
    }

    void refill() {
        for (int blk = 0; blk < 4; ++blk) {
            std::array<uint32_t, 16> init = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                                             key_[0], key_[1], key_[2], key_[3], key_[4], key_[5], key_[6], key_[7],
                                             static_cast<uint32_t>(block_), static_cast<uint32_t>(block_ >> 32), 0u, 0u};
            std::array<uint32_t, 16> s = init;
            for (int dr = 0; dr < 6; ++dr) {  // 12 rounds = 6 column+diagonal double rounds
                quarter(s, 0, 4, 8, 12); quarter(s, 1, 5, 9, 13); quarter(s, 2, 6, 10, 14); quarter(s, 3, 7, 11, 15);
                quarter(s, 0, 5, 10, 15); quarter(s, 1, 6, 11, 12); quarter(s, 2, 7, 8, 13); quarter(s, 3, 4, 9, 14);
            }
            for (int i = 0; i < 16; ++i) buf_[blk * 16 + i] = s[i] + init[i];
            ++block_;
        }
    }

    std::array<uint32_t, 8> key_{};
    uint64_t block_ = 0;
    uint32_t buf_[kBufWords] = {};
    int pos_ = kBufWords;
};

}  // namespace vidx

// rng.hpp -- host-side random stream of the product: what the reference draws from
// rand 0.8.5's StdRng (ChaCha12, rand_chacha 0.3.1) seeded with seed_from_u64
// (rand_core 0.6.4).  Every random decision of k-means (k-means++ picks, batch shuffles,
// reseeds, hierarchy seeds; src/kmeans.rs:31,80,170,240,591) is taken on the host with
// this stream, so a given seed yields the same index as the reference's algorithm.
//
// This is an independent implementation: the test oracle under oracle/ carries its own.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

namespace vidx {

// csrc/chacha_blocks.cpp: eight consecutive ChaCha12 blocks (stream id 0), blocks as SIMD lanes
void chacha12_blocks8(const uint32_t key[8], uint64_t block0, uint32_t* out);

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

    // The first `head` entries of (0..n).shuffle() -- all sample_batch keeps (kmeans.rs:722-726) -- without touching an
    // n-element array at random.  Every draw is made, in order, so the stream ends where the full shuffle leaves it
    // (draw i goes to J[i]: sequential writes).  The last `head` swaps only move entries below `head`; what stands there
    // before them is found by following each of those positions BACK through the earlier swaps: going back, a followed
    // position p is only ever moved by a swap (i - 1, J[i]) with J[i] == p, to i - 1 -- upward, onto a position no other
    // swap still to be visited can name as i - 1 -- so one bitmap test per swap finds the few (about head * ln(n / head))
    // swaps that matter.  The array starts as the identity, so the value is the position the walk ends on.
    void shuffle_head(uint32_t n, uint32_t head, std::vector<uint32_t>& out, std::vector<uint32_t>& scratch) {
        out.resize(std::min(n, head));
        if (n < 2 || head > 1024 || (uint64_t)head * 16 > n) {  // short arrays, long heads (the walk looks a position up linearly)
            scratch.resize(n);
            for (uint32_t i = 0; i < n; i++) scratch[i] = i;
            shuffle(scratch.data(), n);
            std::copy(scratch.begin(), scratch.begin() + out.size(), out.begin());
            return;
        }
        std::vector<uint32_t>& J = scratch;
        J.resize((size_t)n + 1);
        for (uint32_t i = n; i > 1; --i) J[i] = below_u32(i);  // swap (i - 1, J[i]), i = n .. 2
        std::vector<uint64_t> bits(((size_t)n + 63) / 64, 0);
        std::vector<uint32_t> pos(head);
        for (uint32_t t = 0; t < head; t++) {
            pos[t] = t;
            bits[t >> 6] |= 1ull << (t & 63);
        }
        for (uint32_t i = head + 1; i <= n; ++i) {  // backwards in time over the swaps that precede the last `head`
            const uint32_t j = J[i];
            if (!((bits[j >> 6] >> (j & 63)) & 1ull) || j == i - 1) continue;
            uint32_t t = 0;
            while (pos[t] != j) ++t;
            pos[t] = i - 1;
            bits[j >> 6] &= ~(1ull << (j & 63));
            bits[(i - 1) >> 6] |= 1ull << ((i - 1) & 63);
        }
        for (uint32_t i = head; i > 1; --i) std::swap(pos[i - 1], pos[J[i]]);  // the last swaps, forward in time
        std::copy(pos.begin(), pos.end(), out.begin());
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
        return pick_from_prefix(cum, total);
    }
    // The k-means++ draw (kmeans.rs:270-287) squares the distances, sums the weights (`weights.iter().sum()`, the zero
    // test) and lets WeightedIndex::new sum them again for its prefix table.  Both sums are the same sequential chain
    // (0 + w0 = w0 for w0 >= 0), so ONE pass yields the prefix sums and the total: the chain of n dependent f32 adds is the
    // floor of this step and is walked once instead of twice.  Returns the total; n >= 1.
    static float square_prefix(const float* d, size_t n, std::vector<float>& cum) {
        cum.resize(n - 1);
        float run = d[0] * d[0];
        for (size_t i = 1; i < n; ++i) {
            cum[i - 1] = run;
            run += d[i] * d[i];
        }
        return run;
    }
    size_t pick_from_prefix(const std::vector<float>& cum, float total) {
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
    static constexpr int kBufWords = 128;  // eight 16-word blocks per refill (rand_chacha buffers four: the word order is the same)

    static uint32_t rotr(uint32_t v, unsigned r) { r &= 31; return r ? (v >> r) | (v << (32 - r)) : v; }

    void refill() {
        chacha12_blocks8(key_.data(), block_, buf_);
        block_ += 8;
    }

    std::array<uint32_t, 8> key_{};
    uint64_t block_ = 0;
    uint32_t buf_[kBufWords] = {};
    int pos_ = kBufWords;
};

}  // namespace vidx

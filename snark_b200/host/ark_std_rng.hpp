// `ark_std::test_rng()` and `UniformRand` for field elements, for C++ hosts of libb200snark.so.
//
// The reference's entry points take `&mut R: RngCore` (snark/src/lib.rs:43-54); the tests and benches of the reference
// family pass `ark_std::test_rng()`.  None of the generator is in /root/reference (ark-std / rand / rand_chacha / ark-ff
// are un-vendored dependencies), so this restates the published algorithms:
//   * ChaCha with a 64-bit block counter (words 12-13) and a 64-bit stream id (words 14-15): checked against the
//     published ChaCha20/12/8 keystreams through oracle/rng.py (tests/test_oracle_rng.py, tests/test_host_relations.py).
//   * rand 0.8 StdRng = ChaCha12 served through a 64-word BlockRng; test_rng()'s seed; ark-ff's rejection sampling of
//     raw Montgomery limbs: RECALLED, not pinned -- byte parity with arkworks needs golden vectors from a Rust build.
// Host-side only; nothing here runs on the prover hot path.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>

namespace ark_std {

template <int ROUNDS>
class ChaChaRng {
public:
    static constexpr int BUF_WORDS = 64;   // four 16-word blocks per refill
    using Seed = std::array<uint8_t, 32>;

    static ChaChaRng from_seed(const Seed& seed) {
        ChaChaRng r;
        for (int i = 0; i < 8; i++)
            r.key_[i] = uint32_t(seed[4 * i]) | uint32_t(seed[4 * i + 1]) << 8 | uint32_t(seed[4 * i + 2]) << 16 | uint32_t(seed[4 * i + 3]) << 24;
        return r;
    }
    uint32_t next_u32() {
        if (index_ >= BUF_WORDS) generate(0);
        return buf_[index_++];
    }
    // two consecutive words, low half first; with one word left the value straddles the refill
    uint64_t next_u64() {
        const int i = index_;
        if (i < BUF_WORDS - 1) { index_ += 2; return uint64_t(buf_[i + 1]) << 32 | buf_[i]; }
        if (i >= BUF_WORDS) { generate(2); return uint64_t(buf_[1]) << 32 | buf_[0]; }
        const uint64_t lo = buf_[BUF_WORDS - 1];
        generate(1);
        return uint64_t(buf_[0]) << 32 | lo;
    }
    void fill_bytes(uint8_t* out, size_t n) {
        for (size_t k = 0; k < n; k += 4) {
            const uint32_t w = next_u32();
            for (size_t b = 0; b < 4 && k + b < n; b++) out[k + b] = uint8_t(w >> (8 * b));
        }
    }

private:
    static uint32_t rotl(uint32_t x, int n) { return x << n | x >> (32 - n); }
    static void quarter(uint32_t* s, int a, int b, int c, int d) {
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 16);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 12);
        s[a] += s[b]; s[d] = rotl(s[d] ^ s[a], 8);
        s[c] += s[d]; s[b] = rotl(s[b] ^ s[c], 7);
    }
    void block(uint64_t counter, uint32_t* out) const {
        uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};   // "expand 32-byte k"
        for (int i = 0; i < 8; i++) init[4 + i] = key_[i];
        init[12] = uint32_t(counter); init[13] = uint32_t(counter >> 32); init[14] = 0; init[15] = 0;
        uint32_t s[16];
        for (int i = 0; i < 16; i++) s[i] = init[i];
        for (int r = 0; r < ROUNDS / 2; r++) {
            quarter(s, 0, 4, 8, 12); quarter(s, 1, 5, 9, 13); quarter(s, 2, 6, 10, 14); quarter(s, 3, 7, 11, 15);
            quarter(s, 0, 5, 10, 15); quarter(s, 1, 6, 11, 12); quarter(s, 2, 7, 8, 13); quarter(s, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; i++) out[i] = s[i] + init[i];
    }
    void generate(int index) {
        for (int k = 0; k < 4; k++) block(counter_ + k, buf_ + 16 * k);
        counter_ += 4;
        index_ = index;
    }
    uint32_t key_[8] = {};
    uint64_t counter_ = 0;
    uint32_t buf_[BUF_WORDS] = {};
    int index_ = BUF_WORDS;   // empty: the first draw generates
};

using StdRng = ChaChaRng<12>;   // rand 0.8

// ark_std::test_rng(), deterministic branch
inline StdRng test_rng() {
    return StdRng::from_seed({1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0});
}

// `F::rand(rng)` for b2s::Fp<P> (32-bit limbs, Montgomery): ceil(bits/64) u64 draws, least significant first, taken AS
// the Montgomery representation; the top limb is masked to the modulus' bit length; values >= p are redrawn.
template <class F> struct field_params;
template <template <class> class FpT, class P> struct field_params<FpT<P>> { using type = P; };   // Fp<P> -> P

template <class F, class R>
F rand(R& rng) {
    using Params = typename field_params<F>::type;
    constexpr int N64 = F::N / 2;
    static_assert(F::N % 2 == 0, "limb count must be a whole number of u64s");
    constexpr int shave = 64 * N64 - Params::BITS;
    constexpr uint64_t mask = shave == 64 ? 0 : ~uint64_t(0) >> shave;
    for (;;) {
        F x;
        for (int i = 0; i < N64; i++) {
            uint64_t w = rng.next_u64();
            if (i == N64 - 1) w &= mask;
            x.v[2 * i] = uint32_t(w);
            x.v[2 * i + 1] = uint32_t(w >> 32);
        }
        bool less = false;   // x < p ?
        for (int i = F::N - 1; i >= 0; i--) {
            const uint32_t m = Params::mod(i);
            if (x.v[i] != m) { less = x.v[i] < m; break; }
        }
        if (less) return x;
    }
}

}  // namespace ark_std

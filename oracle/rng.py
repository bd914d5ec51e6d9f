"""`ark_std::test_rng()`-compatible random stream and `Fr::rand`, for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

SURVEY.md 8(f) row 3 ("RNG parity").  Nothing of this is in /root/reference (ark-std, rand, rand_chacha and ark-ff are
un-vendored dependencies: Cargo.toml:18-24, no lockfile); it is restated from the published algorithms as recalled:

  * ChaCha block function, 64-bit block counter in words 12-13, 64-bit stream id (0) in words 14-15  -- D. J. Bernstein's
    ChaCha; layout of rand_chacha 0.3.  PINNED here by the published zero-key keystreams for 20, 12 and 8 rounds
    (tests/test_oracle_rng.py).
  * rand 0.8 `StdRng` = ChaCha12; `BlockRng` serves u32 words from a 64-word (4-block) buffer; `next_u64` takes two
    consecutive words, low half first, and straddles a refill when one word is left.            UNPINNED (recalled).
  * ark-std `test_rng()` seed bytes [1,0,0,0, 23,0,0,0, 200,1,0,0, 210,30,0,0, 0 x 16].           UNPINNED (recalled).
  * ark-ff `Fp::rand`: N raw u64 limbs (least significant first) taken AS the Montgomery representation, the top limb
    masked to the modulus' bit length, rejected and redrawn while >= p.                             UNPINNED (recalled).

So: the cipher is certain; how arkworks consumes it can only be confirmed against golden bytes from a Rust build.
"""
from .params import Curve

_MASK = 0xFFFFFFFF
_SIGMA = (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)      # "expand 32-byte k"


def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & _MASK


def _quarter(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & _MASK; s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & _MASK; s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & _MASK; s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & _MASK; s[b] = _rotl(s[b] ^ s[c], 7)


def chacha_block(key_words, counter, stream, rounds):
    """One 64-byte block as 16 little-endian u32 words."""
    init = list(_SIGMA) + list(key_words) + [counter & _MASK, (counter >> 32) & _MASK, stream & _MASK, (stream >> 32) & _MASK]
    s = list(init)
    for _ in range(rounds // 2):
        _quarter(s, 0, 4, 8, 12); _quarter(s, 1, 5, 9, 13); _quarter(s, 2, 6, 10, 14); _quarter(s, 3, 7, 11, 15)
        _quarter(s, 0, 5, 10, 15); _quarter(s, 1, 6, 11, 12); _quarter(s, 2, 7, 8, 13); _quarter(s, 3, 4, 9, 14)
    return [(x + y) & _MASK for x, y in zip(s, init)]


class ChaChaRng:
    """rand_chacha's ChaCha{8,12,20}Rng behind rand_core's BlockRng: 4 blocks per refill."""

    BUF_WORDS = 64

    def __init__(self, seed: bytes, rounds=12):
        assert len(seed) == 32
        self.key = [int.from_bytes(seed[4 * i:4 * i + 4], "little") for i in range(8)]
        self.rounds = rounds
        self.counter = 0
        self.stream = 0
        self.buf = [0] * self.BUF_WORDS
        self.index = self.BUF_WORDS             # empty: the first draw generates

    def _generate(self, index):
        out = []
        for k in range(4):
            out += chacha_block(self.key, self.counter + k, self.stream, self.rounds)
        self.counter = (self.counter + 4) & 0xFFFFFFFFFFFFFFFF
        self.buf = out
        self.index = index

    def next_u32(self):
        if self.index >= self.BUF_WORDS:
            self._generate(0)
        v = self.buf[self.index]
        self.index += 1
        return v

    def next_u64(self):
        n = self.BUF_WORDS
        i = self.index
        if i < n - 1:
            self.index += 2
            return (self.buf[i + 1] << 32) | self.buf[i]
        if i >= n:
            self._generate(2)
            return (self.buf[1] << 32) | self.buf[0]
        lo = self.buf[n - 1]
        self._generate(1)
        return (self.buf[0] << 32) | lo

    def fill_bytes(self, n):
        """Whole u32 words are consumed; a trailing partial word is used from its low bytes up and the rest dropped."""
        out = bytearray()
        while len(out) < n:
            out += self.next_u32().to_bytes(4, "little")
        return bytes(out[:n])


TEST_RNG_SEED = bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16)


def test_rng():
    """ark_std::test_rng() (deterministic branch): StdRng::from_seed(TEST_RNG_SEED)."""
    return ChaChaRng(TEST_RNG_SEED, rounds=12)


test_rng.__test__ = False       # not a pytest test


def field_rand_mont(modulus: int, n_limbs: int, rng) -> int:
    """ark-ff `Fp::rand`: returns the MONTGOMERY representation (the raw accepted limbs) as an int."""
    shave = 64 * n_limbs - modulus.bit_length()
    mask = 0 if shave == 64 else (0xFFFFFFFFFFFFFFFF >> shave)
    while True:
        limbs = [rng.next_u64() for _ in range(n_limbs)]
        limbs[-1] &= mask
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < modulus:
            return v


def fr_rand(curve: Curve, rng) -> int:
    """`Fr::rand(rng)` as the canonical integer: the accepted limbs are a * R, so a = limbs * R^-1 mod r."""
    n = (curve.r.bit_length() + 63) // 64
    mont = field_rand_mont(curve.r, n, rng)
    return mont * pow(1 << (64 * n), -1, curve.r) % curve.r

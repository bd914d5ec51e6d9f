"""oracle/rng.py: the ChaCha core against published keystreams; the BlockRng word/u64 serving rules; Fr::rand rejection
sampling.  (How arkworks consumes the stream is recalled, not pinned -- see the module header.)"""
from oracle import rng as orng
from oracle.params import BLS12_381, BN254


def _keystream(rounds, key=bytes(32), counter=0, stream=0):
    kw = [int.from_bytes(key[4 * i:4 * i + 4], "little") for i in range(8)]
    return b"".join(w.to_bytes(4, "little") for w in orng.chacha_block(kw, counter, stream, rounds)).hex()


def test_chacha_published_vectors():
    # zero key, zero nonce, block 0 -- the classic ChaCha20 / ChaCha12 / ChaCha8 test vectors
    assert _keystream(20) == ("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                              "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
    assert _keystream(12) == ("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
                              "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")
    assert _keystream(8) == ("3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e"
                             "984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42")
    # RFC 7539 section 2.3.2: key 00..1f, block counter 1, nonce 00000009 0000004a 00000000 (words 13..15)
    ks = _keystream(20, bytes(range(32)), counter=1 | (0x09000000 << 32), stream=0x4A000000)
    assert ks.startswith("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e")


def test_block_rng_serving_rules():
    a, b = orng.test_rng(), orng.test_rng()
    words = [a.next_u32() for _ in range(130)]                       # crosses two refills
    assert a.counter == 12 and words[:64] != words[64:128]
    # next_u64 = two consecutive words, low half first
    assert [b.next_u64() for _ in range(3)] == [words[2 * i] | (words[2 * i + 1] << 32) for i in range(3)]
    # one word left in the buffer: the u64 straddles the refill (low = last word, high = first word of the next buffer)
    c = orng.test_rng()
    for _ in range(63):
        c.next_u32()
    assert c.next_u64() == words[63] | (words[64] << 32)
    assert c.next_u32() == words[65]
    # an exhausted buffer: refill, then words 0 and 1
    d = orng.test_rng()
    for _ in range(64):
        d.next_u32()
    assert d.next_u64() == words[64] | (words[65] << 32)
    assert orng.test_rng().fill_bytes(10) == b"".join(w.to_bytes(4, "little") for w in words[:3])[:10]
    # a different seed gives a different stream; the same seed the same
    assert orng.ChaChaRng(bytes(32)).next_u32() == 0x6A9AF49B and orng.test_rng().next_u32() == words[0] != 0x6A9AF49B


def test_field_rand_is_canonical_and_montgomery():
    for curve in (BLS12_381, BN254):
        rng = orng.test_rng()
        seen = [orng.fr_rand(curve, rng) for _ in range(200)]
        assert all(0 <= v < curve.r for v in seen) and len(set(seen)) == 200
        # draws are the raw limbs read as a * R: replaying the stream by hand gives the same elements
        replay = orng.test_rng()
        R = 1 << 256
        mask = (1 << curve.r.bit_length()) - 1
        redone = []
        while len(redone) < 200:
            raw = sum(replay.next_u64() << (64 * i) for i in range(4)) & mask
            if raw < curve.r:
                redone.append(raw * pow(R, -1, curve.r) % curve.r)
        assert redone == seen

    class Fixed:                                                      # rejection: a draw >= p is discarded whole
        def __init__(self, vals):
            self.vals = list(vals)

        def next_u64(self):
            return self.vals.pop(0)

    p = BN254.r
    over = p + 5                                                      # 254 bits, >= p
    limbs = lambda v: [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    top_noise = limbs(7)
    top_noise[3] |= 0xC000000000000000                                # bits 254, 255 are shaved, not rejected
    assert orng.field_rand_mont(p, 4, Fixed(limbs(over) + top_noise)) == 7

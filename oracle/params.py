"""Curve / field parameters for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

This package (`oracle/`) is a CPU restatement of the algorithms on the Groth16
prover hot path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the
product (`snark_b200/`) never does.

PARITY UNPINNED for MSM / NTT / proof values: the reference tree
(arkworks-rs/snark) holds no MSM, NTT, curve or Groth16 code and no golden
vector for them.  The arithmetic lives in the un-vendored
crates ark-ff / ark-ec / ark-poly / ark-groth16 (`^0.5.0`, no lockfile,
`/root/reference/Cargo.toml:17-27`).  The restatement below follows their
published algorithms and is pinned by mathematical
known-answer checks (`tests/test_oracle_*.py`): group order, NTT vs O(n^2)
DFT, MSM vs double-and-add, Groth16 known-trapdoor verification.  The
R1CS matrices x witness seam IS pinned, by the reference's golden matrices
(`relations/src/gr1cs/tests/circuit2.rs:21-43`, `circuit1.rs:28-61`).

All constants here were re-derived or checked numerically (see
tests/test_oracle_fields.py): primality is not re-proved, but generator
order, curve membership and two-adicity are.
"""
from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class Curve:
    name: str
    curve_id: int            # matches B2S_CURVE_* in include/b200snark.h
    p: int                   # base field modulus (Fq)
    r: int                   # scalar field modulus (Fr)
    b: int                   # G1: y^2 = x^3 + b
    b2: Tuple[int, int]      # G2 (over Fq2 = Fq[u]/(u^2+1)): y^2 = x^3 + b2
    g1: Tuple[int, int]
    g2: Tuple[Tuple[int, int], Tuple[int, int]]
    fr_generator: int        # multiplicative generator of Fr (coset offset, ark-ff GENERATOR)
    fr_two_adicity: int
    fq_limbs64: int
    fr_limbs64: int = 4

    @property
    def fr_bits(self):
        return self.r.bit_length()

    @property
    def fr_root_of_unity(self):
        """2^S-th primitive root: GENERATOR^((r-1)/2^S)  (ark-ff TWO_ADIC_ROOT_OF_UNITY)."""
        return pow(self.fr_generator, (self.r - 1) >> self.fr_two_adicity, self.r)

    def omega(self, log_n: int) -> int:
        """Primitive 2^log_n-th root used by ark-poly Radix2EvaluationDomain."""
        assert 0 <= log_n <= self.fr_two_adicity
        return pow(self.fr_root_of_unity, 1 << (self.fr_two_adicity - log_n), self.r)


_BLS_P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
_BLS_R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001

BLS12_381 = Curve(
    name="bls12_381",
    curve_id=0,
    p=_BLS_P,
    r=_BLS_R,
    b=4,
    b2=(4, 4),
    g1=(
        0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
        0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1,
    ),
    g2=(
        (
            0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
            0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e,
        ),
        (
            0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
            0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be,
        ),
    ),
    fr_generator=7,
    fr_two_adicity=32,
    fq_limbs64=6,
)

_BN_P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_BN_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _bn_b2():
    # b2 = 3 / (9 + u) in Fq2
    p = _BN_P
    # (9+u)^-1 = (9 - u) / (81 + 1)
    inv82 = pow(82, -1, p)
    return (3 * 9 * inv82 % p, (-3 * inv82) % p)


BN254 = Curve(
    name="bn254",
    curve_id=1,
    p=_BN_P,
    r=_BN_R,
    b=3,
    b2=_bn_b2(),
    g1=(1, 2),
    g2=(
        (
            10857046999023057135944570762232829481370756359578518086990519993285655852781,
            11559732032986387107991004021392285783925812861821192530917403151452391805634,
        ),
        (
            8495653923123431417604973247489272438418190587263600148770280649306958101930,
            4082367875863433681332203403145435568316851327593401208105741076214120093531,
        ),
    ),
    fr_generator=5,
    fr_two_adicity=28,
    fq_limbs64=4,
)

CURVES = {"bls12_381": BLS12_381, "bn254": BN254, 0: BLS12_381, 1: BN254}

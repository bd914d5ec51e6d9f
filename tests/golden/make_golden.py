#!/usr/bin/env python3
"""Regenerate the committed golden fixtures (tests/golden/*.json).

Two kinds of vectors:
  * reference_matrices.json -- TRANSCRIBED from the reference's own tests (the only golden vectors the reference
    holds for this path): gr1cs/tests/circuit2.rs:21-43, circuit1.rs:28-61, tests/mod.rs:19-33,57-71,138-142.
  * host_vectors.json -- host-side formats, produced by the oracle: the flat LcMap arrays of the finalized circuit2 (the
    input of b2s_r1cs_upload_lcmap), the first words / Fr draws of the restated `ark_std::test_rng()` stream (the ChaCha
    core under it is pinned by published keystreams in tests/test_oracle_rng.py), and circuit2's compressed VerifyingKey.
  * oracle_vectors.json -- produced by the pure-Python oracle (oracle/*.py) with fixed seeds: NTT / coset NTT,
    MSM G1/G2, witness_map and a full Groth16 proof (known trapdoor), compressed proof bytes.  The reference cannot
    be run here (Rust, no toolchain), so these pin the oracle against regressions and give the GPU tests fixed
    known answers; they are NOT outputs of arkworks (parity unpinned, see oracle/params.py).
Integers are written as hex strings.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import groth16 as og  # noqa: E402
from oracle import msm as omsm  # noqa: E402
from oracle import ntt as ontt  # noqa: E402
from oracle import r1cs as orc  # noqa: E402
from oracle import serialize as oser  # noqa: E402
from oracle.ec import groups  # noqa: E402
from oracle.params import BLS12_381, BN254  # noqa: E402

H = lambda x: hex(x)


def pt(P):
    if P is None:
        return None
    if isinstance(P[0], tuple):
        return [[H(P[0][0]), H(P[0][1])], [H(P[1][0]), H(P[1][1])]]
    return [H(P[0]), H(P[1])]


def reference_matrices():
    return {
        "source": "arkworks-rs/snark @ 02fed634, relations/src/gr1cs/tests/",
        "circuit2": {"cite": "circuit2.rs:21-43 (matrices), tests/mod.rs:138-142 (a=1, b=1, c=2)",
                     "assignment": {"a": 1, "b": 1, "c": 2}, "R1CS": orc.CIRCUIT2_GOLDEN},
        "circuit1": {"cite": "circuit1.rs:28-61 (matrices), tests/mod.rs:19-33 (sat), 57-71 (non-sat)",
                     "matrices": orc.CIRCUIT1_GOLDEN, "sat": orc.CIRCUIT1_SAT, "unsat": orc.CIRCUIT1_UNSAT},
    }


def oracle_vectors():
    out = {}
    for curve in (BLS12_381, BN254):
        rng = random.Random(0xB2000000 + curve.curve_id)
        G1, G2 = groups(curve)
        v = {}
        x = [rng.randrange(curve.r) for _ in range(16)]
        v["ntt16"] = {"in": [H(a) for a in x], "fwd": [H(a) for a in ontt.ntt(curve, x)],
                      "inv": [H(a) for a in ontt.ntt(curve, x, inverse=True)],
                      "coset_fwd": [H(a) for a in ontt.coset_ntt(curve, x)], "coset_inv": [H(a) for a in ontt.coset_intt(curve, x)]}
        ks = [rng.randrange(1, curve.r) for _ in range(8)]
        sc = [rng.randrange(curve.r) for _ in range(8)]
        sc[3], sc[5] = 0, curve.r - 1
        b1 = [G1.mul(G1.gen, k) for k in ks]
        b2 = [G2.mul(G2.gen, k) for k in ks]
        v["msm8"] = {"base_logs": [H(k) for k in ks], "scalars": [H(s) for s in sc], "g1_bases": [pt(P) for P in b1],
                     "g2_bases": [pt(P) for P in b2], "g1": pt(omsm.msm_pippenger(G1, b1, sc)), "g2": pt(omsm.msm_pippenger(G2, b2, sc))}
        cs = orc.circuit2(curve, 1, 1, 2)
        cs.finalize()
        mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
        td = og.Trapdoor(*[rng.randrange(1, curve.r) for _ in range(5)])
        rr, ss = rng.randrange(curve.r), rng.randrange(curve.r)
        pk = og.setup(curve, mats, len(inst), len(wit), td)
        A, B, C, h = og.prove(pk, mats, inst, wit, rr, ss)
        assert og.check_in_exponent(pk, (A, B, C), inst, wit, h, rr, ss)
        v["groth16_circuit2"] = {"trapdoor": [H(t) for t in (td.tau, td.alpha, td.beta, td.gamma, td.delta)], "r": H(rr), "s": H(ss),
                                 "h": [H(a) for a in h], "A": pt(A), "B": pt(B), "C": pt(C),
                                 "proof_compressed": oser.proof_compressed(curve, A, B, C).hex(),
                                 "a_query": [pt(P) for P in pk.a_query], "h_query": [pt(P) for P in pk.h_query]}
        out[curve.name] = v
    out["bls12_381"]["generators_compressed"] = {"g1": oser.point_compressed(BLS12_381, 1, groups(BLS12_381)[0].gen).hex(),
                                                  "g2": oser.point_compressed(BLS12_381, 2, groups(BLS12_381)[1].gen).hex(),
                                                  "note": "standard zcash-form encodings of the BLS12-381 generators"}
    return out


def host_vectors():
    from oracle import rng as orng

    out = {}
    for curve in (BLS12_381, BN254):
        cs = orc.circuit2(curve, 1, 1, 2)
        cs.finalize()
        lm = cs.to_lcmap()
        v = {"lcmap_circuit2": {"offsets": lm["offsets"], "vars": [H(x) for x in lm["vars"]], "coeffs": lm["coeffs"],
                                "pool": [H(x) for x in lm["pool"]], "args": [[H(x) for x in a] for a in lm["args"]]}}
        rng = orng.test_rng()
        v["test_rng_fr"] = [H(orng.fr_rand(curve, rng)) for _ in range(4)]
        mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
        pk = og.setup(curve, mats, len(inst), len(wit), og.Trapdoor(3, 5, 7, 11, 13))
        vk = {"alpha_g1": pk.alpha_g1, "beta_g2": pk.beta_g2, "gamma_g2": pk.gamma_g2, "delta_g2": pk.delta_g2, "gamma_abc_g1": pk.gamma_abc_g1}
        v["vk_circuit2_trapdoor_3_5_7_11_13_compressed"] = oser.verifying_key_bytes(curve, vk, True).hex()
        out[curve.name] = v
    rng = orng.test_rng()
    out["test_rng_words"] = [H(rng.next_u32()) for _ in range(16)]
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "host_vectors.json"), "w") as f:
        json.dump(host_vectors(), f, indent=1)
    with open(os.path.join(HERE, "reference_matrices.json"), "w") as f:
        json.dump(reference_matrices(), f, indent=1)
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(oracle_vectors(), f, indent=1)
    print("wrote tests/golden/reference_matrices.json, oracle_vectors.json")

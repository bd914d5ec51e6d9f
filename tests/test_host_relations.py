"""The C++ host mirror of the ark-relations / ark-snark API (snark_b200/host/*.hpp).

CPU part: tests/native/host_relations_test.cpp re-runs the reference's own unit tests (golden circuit2
matrices, satisfiability, variable ordering, LC quirks) against the mirror.  GPU part: the mirror's
Groth16 (`circuit_specific_setup` + `prove`, every group operation through the C ABI) must give the very
proof the Python oracle computes for the same trapdoor and (r, s)."""
import os
import subprocess

import pytest

from oracle import groth16 as og
from oracle import r1cs as orc
from oracle.params import BLS12_381, BN254
from tests.util import unpack_points

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "native", "host_relations_test")
SRC = os.path.join(ROOT, "tests", "native", "host_relations_test.cpp")


def build():
    deps = [SRC] + [os.path.join(ROOT, "snark_b200", "host", f) for f in ("ark_relations.hpp", "ark_snark.hpp")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", EXE, SRC, "-L", os.path.join(ROOT, "snark_b200"), "-lb200snark",
                               "-Wl,-rpath," + os.path.join(ROOT, "snark_b200")])
    return EXE


def test_reference_unit_tests_against_cpp_mirror():
    out = subprocess.run([build(), "cpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bls12_381 cpu tests: ok" in out.stdout and "bn254 cpu tests: ok" in out.stdout and "FAIL" not in out.stdout


def test_cpp_test_rng_matches_oracle_stream():
    """snark_b200/host/ark_std_rng.hpp against oracle/rng.py (itself pinned to the published ChaCha keystreams): raw words
    incl. the u64 that straddles a refill, fill_bytes, and 20 Fr::rand draws per curve."""
    from oracle import rng as orng

    out = subprocess.run([build(), "rng"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = [l.split() for l in out.stdout.splitlines()]
    ref = orng.test_rng()
    assert [int(w, 16) for w in lines[0][1:]] == [ref.next_u32() for _ in range(63 + 2 + 6)]
    assert bytes(int(b, 16) for b in lines[1][1:]) == orng.test_rng().fill_bytes(10)
    for curve, rows in ((BLS12_381, lines[2:22]), (BN254, lines[22:42])):
        ref = orng.test_rng()
        for row in rows:
            assert row[0] == curve.name
            got = sum(int(w, 16) << (32 * i) for i, w in enumerate(row[1:]))
            assert got == orng.field_rand_mont(curve.r, 4, ref)


def test_cpp_lcmap_storage_matches_oracle_export():
    """The C++ mirror keeps linear combinations in the reference's flat LcMap + interner (lc_map.rs:51-56,
    field_interner.rs:13-45); after finalize the arrays -- the input of b2s_r1cs_upload_lcmap -- equal the oracle's
    `to_lcmap()` export, plain and instance-outlined."""
    out = subprocess.run([build(), "lcmap"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    got = {}
    for line in out.stdout.splitlines():
        tag, field, *vals = line.split()
        if field == "pool":
            got.setdefault((tag, field), []).append(sum(int(w, 16) << (32 * i) for i, w in enumerate(vals)))
        else:
            got[(tag, field)] = [int(v) for v in vals]
    for tag, curve, outlined in (("bls12_381", BLS12_381, False), ("bls12_381-outlined", BLS12_381, True), ("bn254", BN254, False)):
        cs = orc.circuit2(curve, 1, 1, 2)
        if outlined:
            cs.set_instance_outliner("R1CS", orc.outline_r1cs)
        cs.finalize()
        lm = cs.to_lcmap()
        assert got[(tag, "offsets")] == lm["offsets"] and got[(tag, "vars")] == lm["vars"] and got[(tag, "coeffs")] == lm["coeffs"], tag
        R = 1 << 256
        assert got[(tag, "pool")] == [v * R % curve.r for v in lm["pool"]], tag                 # Montgomery form
        assert [got[(tag, f"args{k}")] for k in range(3)] == lm["args"], tag


def _parse_words(stdout):
    import numpy as np

    got = {}
    for line in stdout.splitlines():
        tag, *words = line.split()
        got[tag] = np.array([int(w, 16) for w in words], dtype=np.uint32)
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("cid,circuit", [(0, "circuit2"), (1, "dummy")])
def test_cpp_groth16_with_rng_signatures(cid, circuit):
    """`circuit_specific_setup(circuit, rng)` + `prove(pk, circuit, rng)` (snark/src/lib.rs:43-54) on one test_rng()
    stream: alpha, beta, gamma, delta, tau, then r, s -- replayed with oracle/rng.py and checked against the oracle."""
    from oracle import rng as orng

    curve = [BLS12_381, BN254][cid]
    out = subprocess.run([build(), "gpu-rng", str(cid), circuit], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    got = _parse_words(out.stdout)
    rng = orng.test_rng()
    alpha, beta, gamma, delta, tau, rr, ss = (orng.fr_rand(curve, rng) for _ in range(7))
    cs = orc.circuit2(curve, 1, 1, 2) if circuit == "circuit2" else orc.dummy_circuit(curve, 3, 5, 16, 16)
    cs.finalize()
    mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
    pk = og.setup(curve, mats, len(inst), len(wit), og.Trapdoor(tau, alpha, beta, gamma, delta))
    A, B, C, h = og.prove(pk, mats, inst, wit, rr, ss)
    assert og.check_in_exponent(pk, (A, B, C), inst, wit, h, rr, ss)
    assert unpack_points(curve, 1, got["alpha_g1"])[0] == pk.alpha_g1
    assert (unpack_points(curve, 1, got["A"])[0], unpack_points(curve, 2, got["B"])[0], unpack_points(curve, 1, got["C"])[0]) == (A, B, C)


@pytest.mark.gpu
@pytest.mark.parametrize("cid,circuit", [(0, "circuit2"), (0, "dummy"), (1, "circuit2"), (1, "dummy")])
def test_cpp_groth16_matches_oracle(cid, circuit):
    import numpy as np

    curve = [BLS12_381, BN254][cid]
    td_vals = [1234567, 31337, 271828, 314159, 161803]
    rr, ss = 99991, 77773
    out = subprocess.run([build(), "gpu", str(cid), circuit] + [str(v) for v in td_vals + [rr, ss]], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    got = {}
    for line in out.stdout.splitlines():
        tag, *words = line.split()
        got[tag] = np.array([int(w, 16) for w in words], dtype=np.uint32)
    cs = orc.circuit2(curve, 1, 1, 2) if circuit == "circuit2" else orc.dummy_circuit(curve, 3, 5, 16, 16)
    cs.finalize()
    mats, inst, wit = cs.to_matrices(), cs.instance_assignment, cs.witness_assignment
    pk = og.setup(curve, mats, len(inst), len(wit), og.Trapdoor(*td_vals))
    A, B, C, h = og.prove(pk, mats, inst, wit, rr, ss)
    assert og.check_in_exponent(pk, (A, B, C), inst, wit, h, rr, ss)
    assert unpack_points(curve, 1, got["alpha_g1"])[0] == pk.alpha_g1
    assert unpack_points(curve, 1, got["h_query0"])[0] == pk.h_query[0]
    assert (unpack_points(curve, 1, got["A"])[0], unpack_points(curve, 2, got["B"])[0], unpack_points(curve, 1, got["C"])[0]) == (A, B, C)

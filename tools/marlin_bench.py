"""Universal-setup (Marlin-style) prove on one B200: BASELINE config 5's shape (BN254, 2^20 constraints), one JSON line.

A product-chain R1CS (w_{i+2} = w_i * w_{i+1}; one non-zero per row and matrix, columns spread over all variables) of
2^log_n - 3 constraints: |H| = |K| = 2^log_n, SRS of 3 * 2^log_n + 1 powers.  Timed with the device idle before and after
(b2s_sync): universal_setup, index (host arithmetisation + 12 commitments), the per-proof upload (`assign`), and
`prove_assigned` (everything from z on the device to the 10 commitments, 20 evaluations and 2 opening proofs on the host).
The proof is then checked by the independent verifier of oracle/marlin.py (trapdoor form of the opening check)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def chain_circuit(r, log_n, u=3, v=5):
    W = (1 << log_n) - 2
    wit = [u % r, v % r]
    for i in range(W - 2):
        wit.append(wit[-2] * wit[-1] % r)
    A = [[(1, 1)]] + [[(1, 2 + i)] for i in range(W - 2)]
    B = [[(1, 0)]] + [[(1, 3 + i)] for i in range(W - 2)]
    C = [[(1, 2)]] + [[(1, 4 + i)] for i in range(W - 2)]
    return [A, B, C], [1, u % r], wit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--curve", default="bn254", choices=["bn254", "bls12_381"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-verify", action="store_true")
    args = ap.parse_args()
    from snark_b200 import lib as L
    from snark_b200 import marlin as M
    from snark_b200.marlin_gpu import GpuBackend

    gb = GpuBackend(curve=L.BN254 if args.curve == "bn254" else L.BLS12_381, device=0)
    t = time.perf_counter()
    mats, x, w = chain_circuit(gb.r, args.log_n)
    t_circuit = time.perf_counter() - t
    info = M.index_shape(mats, len(x), len(x) + len(w))
    tau = 0x5EED0005B200 % gb.r

    def timed(fn):
        gb.be.sync()
        t0 = time.perf_counter()
        out = fn()
        gb.be.sync()
        return out, (time.perf_counter() - t0) * 1e3

    srs, ms_setup = timed(lambda: gb.setup(info.D + 1, tau))
    (pk, vk), ms_index = timed(lambda: M.index(gb, srs, mats, len(x), len(x) + len(w)))
    z_h, ms_assign = timed(lambda: M.assign(gb, pk, x, w))
    proof, ms_first = timed(lambda: M.prove_assigned(gb, pk, x, z_h, check=True))     # warm-up, with the prover's own checks
    times, launches = [], []
    for _ in range(args.steps):
        l0 = gb.launches
        p2, ms = timed(lambda: M.prove_assigned(gb, pk, x, z_h))
        times.append(ms)
        launches.append(gb.launches - l0)
        assert p2.comms == proof.comms and p2.openings == proof.openings
    verified = None
    if not args.no_verify:
        from oracle import marlin as om
        from oracle.params import BLS12_381, BN254

        curve = BN254 if args.curve == "bn254" else BLS12_381
        verified = bool(om.verify(curve, vk, x, proof, tau=tau))
    print(json.dumps({
        "metric": "marlin_style_proofs_per_sec", "value": 1e3 / min(times), "unit": "proofs/s", "n_gpus": 1,
        "config": {"workload": f"universal-setup (Marlin-style AHP + KZG10) prove, {args.curve}, product-chain R1CS, "
                               f"{info.n_rows} constraints, |H| = 2^{M.log2(info.n)}, |K| = 2^{M.log2(info.m)}, SRS {info.D + 1} powers"},
        "ms_per_proof": min(times), "ms_per_proof_all": times, "ms_first_proof_with_checks": ms_first,
        "ms_universal_setup": ms_setup, "ms_index": ms_index, "ms_assign_upload": ms_assign, "s_circuit_generation_host": t_circuit,
        "gpu_launches_per_proof": launches[-1], "verified_by_oracle": verified, "dtype": "u32 limbs (254/255-bit modular integers)",
        "data": "synthetic"}))
    gb.close()


if __name__ == "__main__":
    main()

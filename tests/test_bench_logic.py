"""bench.py's pure logic on the CPU: synthetic-instance shape, add counts, and the roofline object built from a
per-kernel timing report (a report captured on the B200 is replayed here)."""
import json
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dummy_instance_shape():
    inst = bench.dummy_instance(6)
    assert inst["N"] == 64 and inst["n_rows"] + inst["n_inst"] == 64 and inst["n_inst"] + inst["n_wit"] == 63
    rp, col, co = inst["csr"][0]
    assert rp[0] == 0 and rp[-1] == inst["n_rows"] - 1 and rp[-2] == rp[-1]          # last constraint is empty
    assert set(col.tolist()) == {2} and len(co) == 8 * (inst["n_rows"] - 1)
    assert bench.reference_add_count(1 << 22) == 64880640
    assert bench.msm_window_choice(1 << 24, 192) == (20, 13)


def test_roofline_from_report():
    """A per-kernel report of ONE proof (captured on a B200, profiles/r02_bench_n1_mid.json) replayed: the dominant group is
    the G1 bucket accumulation; its algorithmic bytes are those of all four G1 MSMs of the proof."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n1_mid.json")))
    rep = {k: (1, v) for k, v in d["kernel_ms_per_step"].items()}
    N = 1 << 24
    roof = bench.roofline_from_report(rep, N, 1, 24, 6569.6, "measured")
    assert roof["kernel"].startswith("msm bucket accumulation g1") and roof["traffic"] == bench.NCU_TRAFFIC[("g1", 24, 1)]
    group_ms = sum(v for k, v in d["kernel_ms_per_step"].items() if k in ("msm_accumulate_g1", "msm_ba_p1_g1", "msm_ba_inv_g1", "msm_ba_p2_g1"))
    pairs = 2 * (N - 1) + (N - 3) + (N - 1)
    assert abs(roof["avg_launch_ms"] - group_ms) < 1e-6 and abs(roof["achieved"] - pairs * 128 / (group_ms * 1e-3) / 1e9) < 1e-6
    assert abs(roof["frac"] - roof["achieved"] / 6569.6) < 1e-12 and roof["alu"]["frac"] > 1.0 and set(roof["parts_ms"]) <= set(rep)
    # sharded: a rank consumes 1 / world of the pairs
    roof8 = bench.roofline_from_report(rep, N, 8, 24, 6569.6, "measured")
    assert abs(roof8["algorithmic_bytes_per_launch"] * 8 - roof["algorithmic_bytes_per_launch"]) < 1 and roof8["traffic"] is None
    # a report without MSM kernels still yields a well-formed object
    roof3 = bench.roofline_from_report({"ntt_pass_final<Fr>": (3, 4.5)}, 1 << 24, 1, 24, 6569.6, "measured")
    assert roof3["kernel"] == "ntt_pass_final<Fr>" and roof3["avg_launch_ms"] == 1.5


def test_cpu_sample_scaling_follows_the_reference_add_count():
    """A 2^20 sample stands for 7.2 % of a 2^24 proof, not 1/16: bigger MSMs use fewer window passes per point."""
    assert bench.cpu_scale(24, 24) == 1.0
    assert abs(bench.cpu_scale(20, 24) - (17 * (2**20 + 2**15)) / (15 * (2**24 + 2**18))) < 1e-12
    assert 1 / 16 < bench.cpu_scale(20, 24) < 1 / 13 and 1 / 256 < bench.cpu_scale(16, 24) < 1 / 128


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs first): exactly one JSON line on stdout with the
    contract's keys, `impl: reference`, e2e == value and zero transfer bytes.  Tiny domain so it runs in seconds."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--log-n", "10", "--ref-log-n", "10",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "groth16_proofs_per_sec" and d["unit"] == "proofs/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step"]
    # a sample smaller than the stated configuration: the line carries the step time AS RUN and says it is extrapolated
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--log-n", "12", "--ref-log-n", "10",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][0])
    assert d["extrapolated"] is True and 0 < d["sample_fraction_of_config"] < 1
    assert abs(d["ms_per_step"] * d["value"] / 1e3 - d["sample_fraction_of_config"]) < 1e-9
    assert abs(d["ms_per_step_extrapolated_to_config"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step_extrapolated_to_config"]
    # a non-zero rank of a torchrun launch prints nothing
    quiet = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--log-n", "10", "--ref-log-n", "10",
                            "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root,
                           env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""

"""Why are the batched-affine rounds slow inside prove?  One proof with per-launch timings under a few knobs."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for env in ({"B2S_MSM_AFFINE_ROUNDS": "3"}, {"B2S_MSM_AFFINE_ROUNDS": "3", "B2S_NO_AUX": "1"}):
    e = dict(os.environ); e.update(env); e["B2S_PROFILE_VERBOSE"] = "1"
    print("=====", env, flush=True)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-extras", "--no-cpu"], env=e,
                       capture_output=True, text=True)
    lines = [l for l in p.stderr.splitlines() if l.startswith("[b2s-profile]") and ("ba_round" in l or "accumulate" in l or "horner" in l)]
    print("\n".join(lines[-40:]))
    print(p.stdout[:200])

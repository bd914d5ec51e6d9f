"""Time b2s_msm_g1/g2 at a given size under different tuning knobs (env B2S_MSM_*)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snark_b200

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
groups = [int(g) for g in (sys.argv[2] if len(sys.argv) > 2 else "1").split(",")]
be = snark_b200.Backend(0, 0)
dev = torch.device("cuda", 0)
ext = torch.cuda.ExternalStream(be.stream, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
n = 1 << log_n
def rnd(n):
    t = torch.randint(-(1 << 31), (1 << 31) - 1, (n, 8), dtype=torch.int32, device=dev, generator=g); t[:, 7] &= 0x1FFFFFFF; return t
sc = rnd(n); same = sc[:1].repeat(n, 1).contiguous()
for group in groups:
    bases = torch.empty(n * (be.g1_bytes if group == 1 else be.g2_bytes) // 4, dtype=torch.int32, device=dev)
    be.fixed_base(group, rnd(n), n, mont=False, out=bases); be.sync()
    fn = be.msm_g1 if group == 1 else be.msm_g2
    ref = {}
    for cfg in os.environ.get("PROBE_CFGS", "0:256,1:256,2:256,3:256,2:512,3:512,2:128").split(","):
        r, k = cfg.split(":")
        if r == "auto": os.environ.pop("B2S_MSM_AFFINE_ROUNDS", None)
        else: os.environ["B2S_MSM_AFFINE_ROUNDS"] = r
        for name, s in [kv for kv in (("uniform", sc), ("equal", same)) if kv[0] in os.environ.get("PROBE_KINDS", "uniform,equal").split(",")]:
            out = fn(bases, s, n); be.sync()
            if name not in ref: ref[name] = out
            ok = np.array_equal(ref[name], out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            for _ in range(2): fn(bases, s, n)
            e1.record(ext); be.sync()
            print(f"G{group} 2^{log_n} rounds={r} K={k} {name:8s} {e0.elapsed_time(e1)/2:8.2f} ms  same_result={ok}", flush=True)
            if os.environ.get("PROBE_PROFILE"):
                be.profile(True); fn(bases, s, n); rep = be.profile_report(); be.profile(False)
                top = int(os.environ.get("PROBE_TOP", "9"))
                print(f"    kernel sum {sum(v[1] for v in rep.values()):.2f} ms in {sum(v[0] for v in rep.values())} launches: " +
                      ", ".join(f"{k.split('<')[0]}={v[1]:.2f}" for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:top]), flush=True)
    del bases

"""Turn the raw ncu exports of a GPU session (gpurun_out/) into the committed summaries under profiles/.
usage: python tools/summarize_profiles.py <tag>      (tag = r02)"""
import collections
import csv
import json
import re
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G = "gpurun_out/"
P = "profiles/"
SC = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
TS = {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6, "msecond": 1, "usecond": 1e-3, "second": 1e3, "nsecond": 1e-6}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("b2s::", "")


def launches():
    rows = [r for r in csv.reader(open(G + f"{tag}_launches.csv")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) * TS[r[ui]]
    setup = {k: v for k, v in agg.items() if k.startswith("fixed_base") or k.startswith("pow_table") or k.startswith("msm_precompute")}
    prove = {k: v for k, v in agg.items() if k not in setup}
    total = sum(v[1] for v in prove.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu --no-verify   ({tag}, final code)",
           "# key generation (fixed_base_*) first, then warm-up / timed / e2e / profiled proofs at domain 2^24; cold-cache serialised times -> compare SHARES",
           "#   total_ms share_of_prove count  kernel"]
    for k, v in sorted(prove.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{v[1]:12.3f} {100 * v[1] / total:14.2f} {v[0]:5d}  {k}")
    out.append("\n# setup kernels in the same capture")
    for k, v in sorted(setup.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{v[1]:12.3f} {'':14s} {v[0]:5d}  {k}")
    open(P + f"{tag}_launches_summary.txt", "w").write("\n".join(out) + "\n")
    open(P + f"{tag}_launches.csv", "w").write(open(G + f"{tag}_launches.csv").read())


def traffic():
    rows = [r for r in csv.reader(open(G + f"{tag}_msm_traffic.csv")) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ui, idi = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault((int(r[idi]), short(r[ki])), {})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
    out = ["# one uniform-scalar 2^24-point G1 MSM (= the h-query MSM of a proof): every accumulation kernel launch, in order",
           "# ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active... -k regex:msm_ba_|msm_accumulate  python tools/msm_probe.py 24 1",
           "#  id  kernel                                   dram read GB  write GB   time ms  fmaheavy %   GB/s"]
    tr = tw = tt = 0.0
    for (i, k), m in per.items():
        rd = m["dram__bytes_read.sum"][0] * SC[m["dram__bytes_read.sum"][1]]
        wr = m["dram__bytes_write.sum"][0] * SC[m["dram__bytes_write.sum"][1]]
        t = m["gpu__time_duration.sum"][0] * TS[m["gpu__time_duration.sum"][1]]
        f = m["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]
        out.append(f"{i:5d}  {k[:40]:40s} {rd / 1e9:12.2f} {wr / 1e9:9.2f} {t:9.3f} {f:10.1f} {(rd + wr) / t / 1e6:7.0f}")
        tr += rd; tw += wr; tt += t
    out.append(f"# total of the captured launches: read {tr / 1e9:.2f} GB, write {tw / 1e9:.2f} GB, {tt:.2f} ms (cold-cache, serialised)")
    open(P + f"{tag}_msm_traffic_summary.txt", "w").write("\n".join(out) + "\n")
    return tr + tw


def full(name, outname, header=None):
    rows = list(csv.reader(open(G + name)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "launch__shared_mem_per_block_dynamic", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    out = [header or f"# ncu --set full --import-source on --clock-control none -k regex:msm_ba_p -c 2 python tools/msm_probe.py 24 1   ({tag}, final code; uniform scalars, FIRST round)",
           "# raw page of the report, the metrics the design argues with; stall reasons as shares of the sampled warp states"]
    for r in rows[2:]:
        out.append("\n## " + r[hdr.index("Kernel Name")][:110])
        st = {}
        for h, u, v in zip(hdr, units, r):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v or 0)
            elif h in want:
                out.append(f"  {h:75s} {v} {u}")
        tot = sum(st.values()) or 1
        out.append("  stall reasons: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v / tot > 0.01))
    open(P + outname, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    ntt_hdr = "# ncu --set full --clock-control none -k regex:ntt_pass --launch-skip 3 -c 3 python tools/ntt_probe.py 24   (the three passes of one 2^24 forward NTT, twiddles streamed from the plan's full-size tables)"
    for fn, args in ((launches, ()), (traffic, ()), (full, (f"{tag}_final_ba_g1_raw.csv", f"{tag}_msm_ba_p1_p2_g1_ncu.txt")),
                     (full, (f"{tag}_ntt_full_raw.csv", f"{tag}_ntt_full_ncu.txt", ntt_hdr))):
        try:
            print(fn.__name__, fn(*args))
        except FileNotFoundError as e:
            print("skip", e)

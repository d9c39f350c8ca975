"""Turn the raw ncu artefacts of one round (gpurun_out/prof_<tag>.ncu-rep from `ncu --set full`, and
gpurun_out/launches_<tag>.csv from the `--metrics gpu__time_duration.sum` pass over bench.py) into the tracked
summaries: profiles/<tag>_launches.csv, profiles/<tag>_ncu_summary.md and profiles/roofline_traffic.json.

    python profiles/summarize_ncu.py r1c
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1c"
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = [
    ("gpu__time_duration.sum", "time_ms"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("dram__bytes_read.sum", "dram_read_GB"),
    ("dram__bytes_write.sum", "dram_write_GB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_ncu_peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
]


def short(name):
    n = name.split("(")[0]
    for p in ("void ", "hn::"):
        n = n.replace(p, "")
    return n.split("<")[0]


lines = [f"# ncu summary `{tag}`", ""]
traffic = {}
per_kernel = {}
# samples per MLP launch of profiles/prof_chunk.py 8192: until r1g the fine level was one 8192 x 128 launch; since r1h
# (NerfModel.reuse_coarse_warp) every launch has 8192 x 64 samples: coarse (full program), fine new depths (full program),
# fine inherited depths (trunk-only program)
SPLIT = tag >= "r1h"
rep = os.path.join(OUT, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    what = ("(one 8 192-ray chunk, 0.5 M samples per launch: coarse level, fine level new depths, fine level inherited "
            "depths = trunk-only program, the launch with the smaller numbers). " if SPLIT else
            "(one 8 192-ray chunk: coarse level = 0.5 M samples, fine level = 1 M samples per launch). ")
    lines += ["`ncu --set full --clock-control none --import-source on` of `profiles/prof_chunk.py 8192` " + what +
              "dram % is of ncu's nominal 8 TB/s peak, not of the measured 6.54 TB/s.", "",
              "| kernel | " + " | ".join(k for _, k in KEYS) + " |", "|---|" + "---|" * len(KEYS)]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        vals = []
        for k, _ in KEYS:
            v = d.get(k, "")
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(v)
        lines.append(f"| {short(d['Kernel Name'])} | " + " | ".join(vals) + " |")
        try:
            t = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * 1e9
            key = {"mlp_fwd_kernel": "mlp_fwd", "mlp_dgrad_kernel": "mlp_dgrad", "mlp_wgrad_kernel": "mlp_wgrad"}.get(short(d["Kernel Name"]))
            if key:
                per_kernel.setdefault(key, []).append(t)
        except (KeyError, ValueError):
            pass
    lines.append("")
    for key, ts in per_kernel.items():
        n = 8192 * 64 if SPLIT else 8192 * 128
        traffic[key] = {"dram_bytes_per_launch": max(ts), "samples_in_launch": n, "dram_bytes_per_sample": max(ts) / n,
                        "source": f"profiles/{tag}_ncu_summary.md"}
        if SPLIT:   # the launch with the least traffic is the trunk-only one
            traffic[key + "_trunk"] = {"dram_bytes_per_launch": min(ts), "samples_in_launch": n,
                                       "dram_bytes_per_sample": min(ts) / n, "source": f"profiles/{tag}_ncu_summary.md"}

lc = os.path.join(OUT, f"launches_{tag}.csv")
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration_ns"])
        for r in rows[hi + 1:]:
            if len(r) < len(hdr):
                continue
            name = short(r[ix["Kernel Name"]])[-70:]
            ns = float(r[ix["Metric Value"]])
            w.writerow([r[ix["ID"]], name, r[ix["Grid Size"]], r[ix["Block Size"]], int(ns)])
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += ns / 1e6
    tot = sum(v[1] for v in agg.values())
    lines += ["## launch list", "",
              f"`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` over `bench.py --steps 1 --warmup 3` "
              f"(first 400 launches; per-launch times are cold-cache and serialised: compare shares). Full list: "
              f"`profiles/{tag}_launches.csv`.", "", "| kernel | launches | ms | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
        lines.append(f"| {k} | {v[0]} | {v[1]:.3f} | {v[1] / tot:.3f} |")
    lines.append("")

open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(lines))
if traffic:
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
print("\n".join(lines))

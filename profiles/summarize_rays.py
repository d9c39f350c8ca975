"""profiles/<tag>_rays_ncu_summary.md from gpurun_out/prof_<tag>_rays.ncu-rep (ncu --set full over
`profiles/render_kernel_bw.py 262144 1`: one warm-up and one measured launch of every per-ray kernel)."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1h"
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_rays.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
units = dict(zip(hdr, rows[1]))
keys = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_ncu_peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
out = [f"# ncu summary `{tag}_rays`: the per-ray kernels at 262 144 rays", "",
       "`ncu --set full --clock-control none` over `profiles/render_kernel_bw.py 262144 1` (every second launch of a kernel "
       "is the measured one; the first is its warm-up): coarse sampling, compositing forward / backward at S = 64 and 128 "
       f"(template argument = samples per lane), resampling 64+64 and 64+128. time in {units.get('gpu__time_duration.sum', '')}, "
       f"dram in {units.get('dram__bytes_read.sum', '')}; dram % is of ncu's nominal peak, not of the measured 6.54 TB/s. "
       "Algorithmic bytes and CUDA-event bandwidths: `profiles/README.md`.", "",
       "| kernel | " + " | ".join(k for _, k in keys) + " |", "|---|" + "---|" * len(keys)]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("hn::", "")
    vals = []
    for k, _ in keys:
        v = d.get(k, "")
        try:
            v = f"{float(v):.4g}"
        except ValueError:
            pass
        vals.append(v)
    out.append(f"| {name[:40]} | " + " | ".join(vals) + " |")
open(os.path.join(ROOT, "profiles", f"{tag}_rays_ncu_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))

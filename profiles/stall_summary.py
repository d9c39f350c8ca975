"""Warp-stall summary per kernel from an `ncu --set full --import-source on` report (source page, SASS view):
stall reasons over all samples and the ten most-sampled instructions.  python profiles/stall_summary.py r1g"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1g"
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
out = [f"# warp-stall sampling `{tag}` (ncu source page, SASS)", "",
       "One 8 192-ray chunk (`profiles/prof_chunk.py`); samples cover all 12 warps of a CTA (8 epilogue warps, weight "
       "producer, UMMA issuer, 2 idle), so barrier waits of the idle / feeder warps are part of the totals.", ""]
for kern in ("mlp_fwd", "mlp_dgrad", "mlp_wgrad"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    segs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            segs.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    if not segs:
        continue
    # the largest launch of this kernel (fine level)
    best = None
    for sg in segs:
        hdr = sg["rows"][0]
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in sg["rows"][1:] if len(r) == len(hdr)]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
        if best is None or tot > best[0]:
            best = (tot, hdr, ix, data)
    tot, hdr, ix, data = best
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = sorted(((sum(int(r[ix[h]] or 0) for r in data), h) for h in stalls), reverse=True)
    out += [f"## {kern}_kernel (fine-level launch, {tot} samples)", "",
            "| stall reason | share |", "|---|---|"]
    out += [f"| {h[6:]} | {v / tot:.3f} |" for v, h in agg[:7]]
    out += ["", "| samples | instruction | dominant stall |", "|---|---|---|"]
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:10]:
        st = max(((int(r[ix[h]] or 0), h) for h in stalls))
        out.append(f"| {r[ix['# Samples']]} | `{r[ix['Source']].strip()[:70]}` | {st[1][6:]} |")
    out.append("")
open(os.path.join(ROOT, "profiles", f"{tag}_stalls.md"), "w").write("\n".join(out))
print("\n".join(out))

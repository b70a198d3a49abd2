"""Warp-stall samples of one `ncu --set full --import-source on` report by CUDA source line (top lines).
Usage: python scripts/stalls_by_line.py REPORT.ncu-rep OUT.csv [top_n]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
acc = {}
fname, hdr = None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_s = hdr.index("# Samples")
        i_i = hdr.index("Instructions Executed")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit() or r[2] != "-":
        continue          # rows with an address are SASS lines under their source line
    if not r[i_s].isdigit() or int(r[i_s]) == 0:
        continue
    key = (fname, int(r[0]))
    reasons = sorted(((int(r[i]), h) for i, h in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
    a = acc.setdefault(key, {"samples": 0, "src": r[1].strip(), "instr": 0, "reasons": {}})
    a["samples"] += int(r[i_s])
    a["instr"] += int(r[i_i]) if r[i_i].isdigit() else 0
    for v, h in reasons:
        a["reasons"][h] = a["reasons"].get(h, 0) + v
total = sum(a["samples"] for a in acc.values())
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["samples", "share_pct", "file", "line", "source", "warp_instructions_executed", "top_stall_reasons"])
    for (f, line), a in sorted(acc.items(), key=lambda kv: -kv[1]["samples"])[:top_n]:
        rs = "; ".join(f"{h.replace('stall_', '')} {v}" for h, v in sorted(a["reasons"].items(), key=lambda kv: -kv[1])[:2])
        w.writerow([a["samples"], f"{100.0 * a['samples'] / max(total, 1):.2f}", f, line, a["src"][:140], a["instr"], rs])
print(f"{out}: {len(acc)} source lines with samples, {total} samples")

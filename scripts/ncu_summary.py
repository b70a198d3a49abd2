"""Reduce one `ncu --set full` report (a single kernel launch) to the JSON summary kept under profiles/.
Usage: python scripts/ncu_summary.py REPORT.ncu-rep OUT.json "workload / command line" [key=value ...]
(key=value pairs are copied into the summary, numbers parsed: e.g. updates_per_launch=60000 algorithmic_bytes_per_launch=...)"""
import csv
import io
import json
import subprocess
import sys

rep, out, workload = sys.argv[1], sys.argv[2], sys.argv[3]
extra = dict(a.split("=", 1) for a in sys.argv[4:])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, d = rows[0], rows[1], rows[2]
v = dict(zip(hdr, d))
u = dict(zip(hdr, units))


def num(k, scale=None):
    if k not in v or v[k] == "":
        return None
    x = float(v[k].replace(",", ""))
    unit = u[k]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(unit)
    return x * mult if (scale and mult) else x


s = {"kernel": v["Kernel Name"], "workload": workload, "grid_size": v.get("Grid Size"), "block_size": v.get("Block Size"),
     "gpu_time_ms": num("gpu__time_duration.sum", True) * 1e3,
     "dram_bytes_read": num("dram__bytes_read.sum", True), "dram_bytes_write": num("dram__bytes_write.sum", True),
     "registers_per_thread": num("launch__registers_per_thread"),
     "dynamic_smem_kb": num("launch__shared_mem_per_block_dynamic"),      # reported in Kbyte/block
     "warp_instructions": num("smsp__inst_executed.sum"),
     "issue_slots_busy_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "fp64_pipe_pct_of_peak_active": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
     "lts_sector_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
     "l1tex_sector_hit_rate_pct": num("l1tex__t_sector_hit_rate.pct"),
     "l1tex_lsu_wavefronts_pct_of_peak": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
     "shared_wavefronts_pct_of_peak": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
     "l2_to_l1_read_bytes": num("l1tex__m_xbar2l1tex_read_bytes.sum", True),
     "lts_throughput_pct": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
     "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active")}
s["dram_bytes_per_launch"] = (s["dram_bytes_read"] or 0) + (s["dram_bytes_write"] or 0)
stalls = {}
for h in hdr:
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and v[h]:
        x = float(v[h])
        if x >= 0.1:
            stalls[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(x, 2)
s["warps_stalled_per_issue"] = stalls
for k, val in extra.items():
    try:
        s[k] = float(val) if "." in val or "e" in val else int(val)
    except ValueError:
        s[k] = val
if "algorithmic_bytes_per_launch" in s:
    s["traffic_ratio"] = s["dram_bytes_per_launch"] / s["algorithmic_bytes_per_launch"]
    s["algorithmic_GBps_under_ncu"] = s["algorithmic_bytes_per_launch"] / (s["gpu_time_ms"] * 1e-3) / 1e9
json.dump(s, open(out, "w"), indent=1)
print(json.dumps(s, indent=1))

#!/usr/bin/env python
"""profiles/ncu_traffic.json + a metric summary from one `ncu --set full` capture of the fused
render kernel (tools/run_ncu_fused.sh -> gpurun_out/prof_fused.ncu-rep).  bench.py reads the json
for `roofline.traffic`; the source hash lets it flag the number as stale after a kernel edit.

  python tools/make_traffic_record.py gpurun_out/prof_fused.ncu-rep r02"""
import csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "r02")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
best = None
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if "render_fused" in d.get("Kernel Name", ""):
        best = d
if best is None:
    raise SystemExit("no render_fused launch in the report")
u = dict(zip(hdr, units))
def val(name):
    v = float(best[name].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[name], 1)
    return v * scale
dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
sources = ["cr-nerf-pytorch_b200/csrc/nerf_mlp.cu", "cr-nerf-pytorch_b200/csrc/nerf_layout.h", "cr-nerf-pytorch_b200/csrc/ptx.cuh"]
h = hashlib.sha256()
for rel in sources:
    h.update(open(os.path.join(ROOT, rel), "rb").read())
rec = {"kernel": best["Kernel Name"], "grid": best.get("Grid Size"), "dram_bytes_per_launch": dram,
       "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
       "capture": f"ncu --set full --clock-control none, profiles/{tag}_ncu_fused_fine_pass_metrics.txt",
       "sources": sources, "sources_sha256": h.hexdigest()}
json.dump(rec, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
keys = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active")
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_fused_fine_pass_metrics.txt"), "w") as f:
    f.write(f"# {best['Kernel Name']}  grid {best.get('Grid Size')} block {best.get('Block Size')}\n# from {os.path.basename(rep)} (ncu --set full --clock-control none)\n")
    for k in keys:
        if k in best:
            f.write(f"{k} [{u[k]}] = {best[k]}\n")
print(json.dumps(rec, indent=1))

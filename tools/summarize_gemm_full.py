#!/usr/bin/env python
"""Merge gpurun_out/r2_gemm_full.ncu-rep (tools/capture_gemm_full.sh) with the launch list of the same evaluation and write
profiles/r2_gemm_full.summary.json / .txt: per launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) against the
algorithmic bytes, tensor-pipe and DRAM utilisation.  bench.py reads the JSON for `roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2_gemm_full.ncu-rep")
lst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "r2_gemm_launch_list.json")
csv_path = os.path.splitext(rep)[0] + "_raw.csv"          # exported on the GPU box by tools/capture_gemm_full.sh
if os.path.exists(csv_path):
    raw = open(csv_path).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
raw = raw[raw.index('"ID"'):] if '"ID"' in raw else raw
rows = list(csv.DictReader(io.StringIO(raw)))[1:]          # first data row holds the units
launches = [e for e in json.load(open(lst)) if e["flops"] > 0]      # conv_in is not a gemm_tc_kernel launch: ncu -k regex:gemm_tc skipped it


def unit_scale(col):
    units = list(csv.DictReader(io.StringIO(raw)))[0]
    u = units.get(col, "")
    return {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)


sb_r, sb_w, st = unit_scale("dram__bytes_read.sum"), unit_scale("dram__bytes_write.sum"), unit_scale("gpu__time_duration.sum")
out, txt = [], []
for i, r in enumerate(rows):
    L = launches[i] if i < len(launches) else dict(name="?", flops=0.0, bytes=0.0)
    dram = float(r["dram__bytes_read.sum"]) * sb_r + float(r["dram__bytes_write.sum"]) * sb_w
    e = dict(i=i, name=L["name"], kernel=r["Kernel Name"].split("(")[0].replace("void seer::", ""), ncu_us=float(r["gpu__time_duration.sum"]) * st,
             dram_bytes=dram, algorithmic_bytes=L["bytes"], ratio=dram / L["bytes"] if L["bytes"] else None,
             tensor_pipe_pct=float(r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]),
             dram_pct=float(r["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]))
    out.append(e)
    txt.append(f"{i:3d} {e['name']:40s} {e['kernel']:28s} {e['ncu_us']:8.1f} us  dram {dram / 1e6:8.1f} MB  algo {L['bytes'] / 1e6:8.1f} MB  "
               f"x{(e['ratio'] or 0):.2f}  tensor {e['tensor_pipe_pct']:5.1f}%  dram {e['dram_pct']:5.1f}%")
mean_dram = sum(e["dram_bytes"] for e in out) / len(out)
mean_algo = sum(e["algorithmic_bytes"] for e in out) / len(out)
summary = dict(what="first %d gemm_tc_kernel launches of one eager UNet evaluation at the bench shape (level-0 and level-1 down blocks)" % len(out),
               launches=len(out), traffic_bytes_per_launch_mean=mean_dram, algorithmic_bytes_per_launch_mean=mean_algo,
               traffic_over_algorithmic=mean_dram / mean_algo, per_launch=out)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(summary, open(os.path.join(ROOT, "profiles", "r2_gemm_full.summary.json"), "w"), indent=1)
head = (f"ncu --set full, {len(out)} gemm_tc_kernel launches (tools/capture_gemm_full.sh): mean DRAM traffic {mean_dram / 1e6:.1f} MB per launch vs "
        f"{mean_algo / 1e6:.1f} MB algorithmic (x{mean_dram / mean_algo:.2f}); ncu times are cold-cache / serialised\n")
open(os.path.join(ROOT, "profiles", "r2_gemm_full.summary.txt"), "w").write(head + "\n".join(txt) + "\n")
print(head)

#!/usr/bin/env python
"""Turn the reports of tools/ncu_capture.sh (gpurun_out/<tag>_{strict,fast}.ncu-rep) into the tracked summaries under
profiles/ and re-stamp profiles/roofline_traffic.json (DRAM bytes, instruction counts, sha of the CUDA sources).

usage: python tools/ncu_summarise.py <tag>        (run in the build container: needs ncu -i, cuobjdump, nvdisasm)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tag = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: [v, u] for h, u, v in zip(hdr, units, vals)}


def to_number(v, unit):
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
    return x * scale.get(unit, 1.0)


prof = {}
for mode in ("strict", "fast"):
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{mode}.ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep)
        continue
    m = raw_metrics(rep)
    keep = {k: m[k] for k in KEYS if k in m}
    json.dump(keep, open(os.path.join(ROOT, "profiles", f"{tag}_fused_step_{mode}_ncu_metrics.json"), "w"), indent=1)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", f"{tag}_fused_step_{mode}_ncu_details.txt"), "w").write(det)
    op = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_opmix.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", f"{tag}_fused_step_{mode}_opmix.txt"), "w").write(op)
    pat = "k_fused_stepILi2ELb1ELi2ELi0ELi1E" if mode == "strict" else "k_fused_stepILi2ELb1ELi2ELi1ELi0E"
    hl = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot_lines.py"), rep,
                         os.path.join(ROOT, "euler2d_kokkos_b200", "libeuler2d_b200.so"), pat, "45"], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", f"{tag}_fused_step_{mode}_hot_lines.txt"), "w").write(hl)
    key = "k_fused_step_8192x8192" if mode == "strict" else "k_fused_step_fast_8192x8192"
    rd = to_number(*m["dram__bytes_read.sum"])
    wr = to_number(*m["dram__bytes_write.sum"])
    prof[key] = rd + wr
    prof[key + "_dram_read"] = rd
    prof[key + "_dram_write"] = wr
    prof[key + "_warp_inst"] = float(m["smsp__inst_executed.sum"][0].replace(",", ""))
    fp64 = 0
    for line in op.splitlines():
        if line.startswith("FP64-pipe:"):
            fp64 = int(line.split()[1])
    prof[key + "_fp64_warp_inst"] = fp64
    prof[key + "_ncu_ms"] = to_number(*m["gpu__time_duration.sum"]) * 1e3

from bench import csrc_sha  # noqa: E402

old = {}
try:
    old = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
except Exception:
    pass
prof["fp64_peak_warp_inst_per_clk_per_smsp"] = old.get("fp64_peak_warp_inst_per_clk_per_smsp", 0.476)
prof["fp64_peak_source"] = old.get("fp64_peak_source", "")
prof["csrc_sha"] = csrc_sha()
prof["_source"] = (f"tools/ncu_capture.sh {tag} (ncu --set full --clock-control none -k regex:k_fused_step -s 6 -c 1 python "
                   f"tools/quick_perf.py four_quadrant 8192 8192 2 strict|fast) summarised by tools/ncu_summarise.py; "
                   f"profiles/{tag}_fused_step_*")
json.dump(prof, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
print(json.dumps(prof, indent=1))

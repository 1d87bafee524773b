#!/usr/bin/env python3
"""Turn `ncu -i X.ncu-rep --page raw --csv` into the markdown + json summaries kept under profiles/.
usage: ncu_summary.py raw.csv out.md [traffic.json]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__warps_eligible.avg.per_cycle_active"]
out = ["| kernel | metric | unit | value |", "|---|---|---|---|"]
traffic = {}
for vals in rows[2:]:
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    name = d.get("Kernel Name", ("", "?"))[1][:60]
    for k in KEYS:
        if k in d:
            out.append(f"| {name} | {k} | {d[k][0]} | {d[k][1]} |")
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(d[h][1] or 0) > 0.05:
            out.append(f"| {name} | stall:{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | warps/issue | {d[h][1]} |")
    def mb(k):
        u, v = d.get(k, ("", "0"))
        f = float(v or 0)
        return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    if "k_aggregate" in name:
        traffic["k_aggregate_dram_bytes_per_launch"] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
        traffic["k_aggregate_ms_under_ncu"] = float(d["gpu__time_duration.sum"][1])
        traffic["source"] = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one " + name.split("(")[0].strip() + " launch at C2 on one GPU"
open(sys.argv[2], "w").write("\n".join(out) + "\n")
if len(sys.argv) > 3:
    json.dump(traffic, open(sys.argv[3], "w"), indent=1)
print("\n".join(out[:12]))

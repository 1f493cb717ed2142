#!/usr/bin/env python
"""Summarise an ncu report for profiles/: one column per captured kernel, the metrics the roofline rests on.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv profiles/rN_ncu_summary.csv [profiles/rN_kernel_traffic.json]
"""
import csv
import json
import sys

METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
SHORT = {"k_sweepA": "sweepA", "k_diss": "dissipation", "k_sweepBD": "sweepB", "k_sweepB": "sweepB",
         "k_adjoint1v2": "adjoint1", "k_adjoint1": "adjoint1",
         "k_adjoint2": "adjoint2"}


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
    return float(value.replace(",", "")) * scale


def main():
    """argv: raw.csv[,raw2.csv,...] out_summary.csv [out_traffic.json] -- several raw pages (one per capture) are
    merged column-wise; every cell carries its own unit because ncu scales units per report."""
    cols = []          # (kernel name, {metric: (value, unit)})
    for path in sys.argv[1].split(","):
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        for d in data:
            cols.append((d[hdr.index("Kernel Name")], {h: (v, u) for h, v, u in zip(hdr, d, units)}))
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric"] + [n for n, _ in cols])
        for m in METRICS:
            if any(m in c for _, c in cols):
                w.writerow([m] + [(c[m][0] + (" " + c[m][1] if c[m][1] else "")) if m in c else "" for _, c in cols])
    if len(sys.argv) > 3:
        traffic = {}
        tscale = {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}
        for n, c in cols:
            key = next((v for k, v in SHORT.items() if k + "<" in n or k + "(" in n), None)
            if key is None:
                continue
            e = traffic.setdefault(key, {"dram_bytes_per_launch": [], "ncu_ms": []})
            e["dram_bytes_per_launch"].append(to_bytes(*c["dram__bytes_read.sum"]) + to_bytes(*c["dram__bytes_write.sum"]))
            t = c["gpu__time_duration.sum"]
            e["ncu_ms"].append(float(t[0].replace(",", "")) * tscale.get(t[1], 1.0))
        out = {k: {"dram_bytes_per_launch": sum(v["dram_bytes_per_launch"]) / len(v["dram_bytes_per_launch"]),
                   "ncu_ms": sum(v["ncu_ms"]) / len(v["ncu_ms"]), "launches_captured": len(v["ncu_ms"])}
               for k, v in traffic.items()}
        out["_source"] = "ncu --set full --clock-control none, one capture per kernel of `python bench.py --steps 1 --warmup 0`"
        json.dump(out, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()

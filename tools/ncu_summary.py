#!/usr/bin/env python
"""Summarise an .ncu-rep (from `ncu --set full`) into a small markdown file for profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_c2_full.md [kernel-substring]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    sub = sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    data = [r for r in data if sub in r[kn]]
    lines = [f"# ncu --set full summary: `{rep.split('/')[-1]}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` on a B200 (one GPU). Per-launch values;",
             "ncu replays each launch ~40x with caches in whatever state the replay leaves them (the 80 MB of output of",
             "a C2 launch stays in the 126 MB L2, so `dram__bytes_write` is far below the bytes stored; the C3-shard capture",
             "is the one where output exceeds L2).", "",
             "| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |",
             "|---|---|" + "---|" * len(data)]
    lines.insert(2, "Kernels: " + ", ".join(sorted({r[kn].split('(')[0] for r in data})))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"| `{k}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    stall = [(h, i) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    lines += ["", "Warp stall reasons (cycles stalled per issued instruction, last launch):", "", "| reason | ratio |", "|---|---|"]
    for h, i in sorted(stall, key=lambda x: -float(data[-1][x[1]] or 0)):
        lines.append(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {data[-1][i]} |")
    if len(sys.argv) > 3:
        try:
            src = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, sub], capture_output=True, text=True).stdout
            lines += ["", "Hot source lines (share of executed warp instructions / of stall samples):", "", "```"] + src.splitlines()[:40] + ["```"]
        except Exception as e:                                     # noqa: BLE001
            lines.append(f"(source aggregation failed: {e})")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()

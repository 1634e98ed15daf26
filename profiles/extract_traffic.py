#!/usr/bin/env python
"""Turn `ncu --set full` captures (.ncu-rep, brought back in gpurun_out/) into tracked evidence:

    python profiles/extract_traffic.py NAME=gpurun_out/file.ncu-rep [NAME=...]

For every report: `ncu -i file --page raw --csv` -> profiles/ncu_NAME.csv (one row per profiled launch, the columns the
judge reads: duration, dram bytes, tensor / XU / FMA / ALU pipe, issue slots, L2 hit, registers, stall reasons), and
profiles/kernel_traffic.json gets, per kernel family bench.py ranks, the per-launch dram__bytes_read.sum + dram__bytes_write.sum
of the LONGEST launch of that kernel in the report (bench.py's `roofline.traffic` reads that file -- it holds no literal).
"""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
COLS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
# kernel-name substring -> the name bench.py's roofline uses
FAMILY = [("flash_attn", "flash_attn"), ("ffn_fused", "ffn_fused"), ("rel_attn", "rel_attn"), ("istft", "istft"),
          ("source_stft", "source_stft"), ("nsf_source", "nsf_source")]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    traffic_path = os.path.join(HERE, "kernel_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, units, data = rows[0], rows[1], rows[2:]
        idx = [head.index(c) for c in COLS if c in head]
        with open(os.path.join(HERE, f"ncu_{name}.csv"), "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow([head[i] for i in idx])
            w.writerow([units[i] for i in idx])
            for r in data:
                w.writerow([r[i] for i in idx])
        kn, dur = head.index("Kernel Name"), head.index("gpu__time_duration.sum")
        rd, wr = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
        for sub, fam in FAMILY:
            cand = [r for r in data if sub in r[kn]]
            if not cand:
                continue
            r = max(cand, key=lambda r: float(r[dur].replace(",", "")))
            traffic[fam] = {"dram_bytes_per_launch": to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr]),
                            "dram_read_bytes": to_bytes(r[rd], units[rd]), "dram_write_bytes": to_bytes(r[wr], units[wr]),
                            "launch_us_under_ncu": float(r[dur].replace(",", "")), "kernel": r[kn].split("(")[0],
                            "source": f"profiles/ncu_{name}.csv (ncu --set full --clock-control none, longest launch of the kernel)"}
        print(name, len(data), "launches ->", f"profiles/ncu_{name}.csv")
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

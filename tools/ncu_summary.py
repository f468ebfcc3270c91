#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into one row per profiled launch:
duration, DRAM bytes, throughput fractions, occupancy, issue/stall picture, smem conflicts.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/rN_xxx.txt]"""
import csv
import io
import subprocess
import sys

UNIT = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
WANT = [
    ("gpu__time_duration.sum", "dur_us", 1e6),       # scale applies after conversion to base units (s, byte)
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__occupancy_limit_registers", "lim_reg", 1),
    ("launch__occupancy_limit_shared_mem", "lim_smem", 1),
    ("sm__inst_executed.sum", "inst_M", 1e-6),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct", 1),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct", 1),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct", 1),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct", 1),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf_M", 1e-6),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wave_M", 1e-6),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long", 1),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar", 1),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short", 1),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio", 1),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg", 1),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math", 1),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait", 1),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel", 1),
]


def short(name):
    import re
    m = re.search(r"(frame_kernel\w*)<.*?(Cfft|RfftFwd|RfftInv|Mfcc)\w*<.*?Arith(\w+?), *(\d+), *(\d+), *(\d+)", name)
    if m:
        return f"{m.group(1)}:{m.group(2)}:{m.group(3)}:N{m.group(4)}:T{m.group(5)}:F{m.group(6)}"
    return name[:60]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kcol = col["Kernel Name"]
    names = [n for n, _, _ in WANT if n in col]
    print("# " + rep)
    print("kernel grid block " + " ".join(lbl for n, lbl, _ in WANT if n in col))
    for r in rows[2:]:
        vals = []
        for n, lbl, sc in WANT:
            if n not in col:
                continue
            v = r[col[n]].replace(",", "")
            try:
                vals.append(f"{float(v) * UNIT.get(units[col[n]], 1.0) * sc:.4g}")
            except ValueError:
                vals.append(v or "-")
        print(short(r[kcol]), r[col["Grid Size"]].replace(" ", ""), r[col["Block Size"]].replace(" ", ""), " ".join(vals))


if __name__ == "__main__":
    main()

"""Condense `ncu -i X.ncu-rep --page raw --csv` output into a small markdown table (one column per kernel).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > prof_raw.csv ; python tools/ncu_summary.py prof_raw.csv > profiles/NAME.md
"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid (CTAs)"),
    ("launch__block_size", "block (threads)"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (shared memory)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of 64/SM)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots used (%)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU/SFU) pipe (%)"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active (%)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (%)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (%)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts (%)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"
STALL_SUFFIX = "_per_issue_active.ratio"


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")].replace("void ", "").replace("tcelbo::", "")[:48] for r in data]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in KEYS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for r in data:
            try:
                v = float(r[i])
                vals.append(f"{v:,.3f}".rstrip("0").rstrip(".") + (f" {units[i]}" if units[i] and units[i] != "%" else ""))
            except ValueError:
                vals.append(r[i])
        print(f"| {label} | " + " | ".join(vals) + " |")
    stall_cols = [(i, h[len(STALL_PREFIX):-len(STALL_SUFFIX)]) for i, h in enumerate(hdr)
                  if h.startswith(STALL_PREFIX) and h.endswith(STALL_SUFFIX)]
    for i, name in stall_cols:
        vals = [float(r[i]) for r in data]
        if max(vals) >= 0.2 and name != "selected":
            print(f"| stall: {name} (warps per issued instruction) | " + " | ".join(f"{v:.2f}" for v in vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])

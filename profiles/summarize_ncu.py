"""Condenses an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py cite.

    python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep > profiles/r1_x_ncu_summary.txt
"""
import csv
import io
import re
import subprocess
import sys

KEEP = [
    r"^gpu__time_duration\.sum$", r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_sectors_srcunit_tex_op_(read|write)\.sum$", r"^lts__t_sector_hit_rate\.pct$",
    r"^l1tex__t_sectors_pipe_lsu_mem_global_op_(ld|st)\.sum$", r"^l1tex__t_requests_pipe_lsu_mem_global_op_(ld|st)\.sum$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$",
    r"^smsp__inst_executed\.sum$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__warps_eligible\.avg\.per_cycle_active$",
    r"^sm__inst_executed_pipe_(alu|fma|lsu|adu|xu)\.avg\.pct_of_peak_sustained_active$",
    r"^launch__(registers_per_thread|grid_size|block_size|waves_per_multiprocessor|occupancy_limit_(registers|shared_mem|warps)|shared_mem_per_block_dynamic)$",
    r"^smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio$",
    r"^sm__cycles_elapsed\.max$", r"^smsp__cycles_active\.avg$",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("kernel:", name)
        for i, h in enumerate(hdr):
            if any(re.search(k, h) for k in KEEP):
                v = row[i]
                if "stalled" in h:
                    try:
                        if float(v) < 0.2:
                            continue
                    except ValueError:
                        pass
                print(f"  {h} [{units[i]}] = {v}")


if __name__ == "__main__":
    main(sys.argv[1])

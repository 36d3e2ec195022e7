#!/usr/bin/env python3
"""Turn the ncu reports of a GPU-box call (gpurun_out/*.ncu-rep, scratch) into what profiles/ keeps under version control:
  profiles/<tag>_<name>_raw.csv      `ncu --page raw --csv` of the captured launch (every metric of --set full)
  profiles/<tag>_<name>_summary.txt  the handful the roofline / issue-slot discussion quotes
and, with --sass, an excerpt of the shipped trace_kernel SASS showing the TMA bulk copies and mbarrier waits.
Usage: python tools/ncu_export.py TAG name=report.ncu-rep [name=report.ncu-rep ...] [--sass]"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def main():
    tag = sys.argv[1]
    for spec in sys.argv[2:]:
        if spec == "--sass":
            so = os.path.join(ROOT, "phosphorus_mk2_b200", "lib", "libphos_cuda.so")
            sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
            out, fn, keep = [], None, 0
            for i, line in enumerate(sass):
                if "Function :" in line:
                    fn = line.strip()
                if fn and "trace_kernel" in fn and any(k in line for k in ("UBLKCP", "SYNCS", "LDG.E.128", "MATCH", "STL", "LDL")):
                    out.append(f"{fn}\n    " + "\n    ".join(l.rstrip() for l in sass[max(0, i - 1):i + 2]))
            path = os.path.join(ROOT, "profiles", f"{tag}_sass_trace_kernel.txt")
            with open(path, "w") as f:
                f.write("# cuobjdump -sass phosphorus_mk2_b200/lib/libphos_cuda.so, trace_kernel instantiations: the TMA bulk copies (UBLKCP),\n"
                        "# mbarrier operations (SYNCS), 128-bit loads (LDG.E.128), MATCH (tail work sharing) and any local-memory access (STL / LDL)\n")
                f.write("\n".join(out) + "\n")
            print("wrote", path, len(out), "sites")
            continue
        name, rep = spec.split("=", 1)
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        raw = raw[raw.index('"ID"'):] if '"ID"' in raw else raw
        open(os.path.join(ROOT, "profiles", f"{tag}_{name}_raw.csv"), "w").write(raw)
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        with open(os.path.join(ROOT, "profiles", f"{tag}_{name}_summary.txt"), "w") as f:
            for r in rows[2:]:
                d = dict(zip(hdr, r))
                f.write(f"# {d.get('Kernel Name', '?')}  (ncu --set full --clock-control none, report {os.path.basename(rep)})\n")
                for k in KEEP:
                    if k in d:
                        f.write(f"{k:90s} {d[k]:>18s} {units[hdr.index(k)]}\n")
        print("wrote", name)


if __name__ == "__main__":
    main()

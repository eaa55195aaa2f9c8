#!/usr/bin/env python
"""Condense one `ncu --set full` report into the JSON kept under profiles/:

    python tools/ncu_summary.py gpurun_out/f_full.ncu-rep profiles/rNN_ncu_summary.json "command line of the capture" [blocks_per_launch]

Reads `ncu -i <rep> --page raw --csv` (one kernel launch per row) and keeps the metrics DESIGN.md / profiles/README.md quote.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed launch__registers_per_thread smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
sass__inst_executed_local_loads sass__inst_executed_local_stores sm__cycles_elapsed.avg sm__warps_active.avg.per_cycle_active
sm__throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct""".split()


def main():
    rep, out, command = sys.argv[1:4]
    blocks = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
    text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    names, units, vals = rows[0], rows[1], rows[2]
    metrics = {}
    for n, u, v in zip(names, units, vals):
        if n in KEEP or n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued"):
            try:
                metrics[n] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                pass
    kernel = vals[names.index("Kernel Name")]
    json.dump({"kernel": kernel, "command": command, "blocks_per_launch": blocks, "metrics": metrics}, open(out, "w"), indent=1)
    for n in KEEP[:12]:
        if n in metrics:
            print(n, metrics[n]["value"], metrics[n]["unit"])


if __name__ == "__main__":
    main()

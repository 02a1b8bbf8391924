"""Key metrics of an ncu report (raw page) as a short text summary.  usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum\t", "lts__t_bytes.sum ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "smsp__inst_executed.sum ", "sm__cycles_elapsed.avg ", "smsp__average_warps_issue_stalled", "sm__throughput.avg.pct",
        "local_load", "local_store", "smsp__inst_executed_op_local"]
for vals in rows[2:]:
    name = dict(zip(hdr, vals)).get("Kernel Name", "?")
    print("==", name[:150])
    for h, u, v in zip(hdr, units, vals):
        hh = h + " "
        if any(k in hh for k in KEYS) or h in ("dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
                                                  "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"):
            if "stalled" in h and "per_issue_active" not in h:
                continue
            print(f"  {h} [{u}] = {v}")

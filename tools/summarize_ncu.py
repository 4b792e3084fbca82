"""One .ncu-rep (ncu --set full) -> the metrics DESIGN.md / bench.py quote, as text.
Usage: python tools/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/rNN_x.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe", "sm__inst_executed_pipe_tensor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")])
    for i, h in enumerate(hdr):
        if any(h.startswith(w) or w in h for w in WANT) and r[i] != "":
            print(f"  {h:75s} {r[i]:>18s} {units[i]}")

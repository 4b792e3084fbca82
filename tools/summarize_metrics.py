"""ncu --metrics ... --csv log -> one line per distinct launch shape (kernel, grid, bytes), averaged, as text.
Usage: python tools/summarize_metrics.py gpurun_out/x_metrics.csv > profiles/rNN_x_metrics.txt"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

SHORT = OrderedDict([
    ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("lts__t_bytes.sum", "l2MB"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("sm__inst_executed_pipe_xu.sum", "xuMinst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smemMwave"), ("smsp__inst_executed.sum", "Minst"),
    ("sm__cycles_elapsed.avg", "cycles")])


def conv(v, unit, short):
    v = float(v.replace(",", ""))
    if short == "us":
        return v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    if short.endswith("MB"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        return v * mult / 1e6
    if short in ("xuMinst", "smemMwave", "Minst"):
        return v / 1e6
    return v


with open(sys.argv[1], newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ii, ki, mi, ui, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
launch = OrderedDict()
for r in rd:
    if len(r) <= vi or r[mi] not in SHORT:
        continue
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki]))
    d = launch.setdefault(r[ii], {"kernel": name})
    d[SHORT[r[mi]]] = conv(r[vi], r[ui], SHORT[r[mi]])
groups = defaultdict(list)
for d in launch.values():
    key = (d["kernel"], int(d.get("grid", 0)), round(d.get("rdMB", 0) + d.get("wrMB", 0), -1), round(d.get("us", 0), -1 if d.get("us", 0) > 100 else 0))
    groups[key].append(d)
cols = [c for c in SHORT.values() if c != "grid"]
print(f"{'kernel':44s} {'n':>3s} {'grid':>5s} " + " ".join(f"{c:>9s}" for c in cols))
for (k, grid, _, _), ds in sorted(groups.items(), key=lambda kv: -sum(d.get("us", 0) for d in kv[1])):
    avg = {c: sum(d.get(c, 0) for d in ds) / len(ds) for c in cols}
    print(f"{k[-44:]:44s} {len(ds):3d} {grid:5d} " + " ".join(f"{avg[c]:9.1f}" for c in cols))

"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel count / total time / share, as text.

Usage: python tools/summarize_launches.py gpurun_out/launches.csv [--vae-chunks 64] > profiles/rNN_launches_summary.txt

ncu serialises and cold-starts every launch (about 150 ms of profiler overhead per launch on this box), so a full
config-3 step (7.5 K launches) does not fit a GPU call.  With --vae-chunks N the list may stop inside the VAE decode:
the launches before the first `vae_stem_kernel` (both DiT forwards, fan-out, x0) are taken as they are, the COMPLETE
decoder chunks seen (a chunk = the launches from one `vae_stem_kernel` to the next; all chunks of a step are
identical work on 128 tiles) are averaged, and the step is pre-VAE + N x that average.  The rule / select kernels after
the decode (< 0.1 % of a step) are then missing from the table.
"""
import csv
import re
import sys
from collections import defaultdict


def load(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    out = []
    for r in rd:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name)
        out.append((name, us))
    return out


def main():
    path = sys.argv[1]
    n_chunks = int(sys.argv[sys.argv.index("--vae-chunks") + 1]) if "--vae-chunks" in sys.argv else 0
    rows = load(path)
    agg = defaultdict(lambda: [0.0, 0.0])
    note = ""
    if n_chunks:
        stems = [i for i, (n, _) in enumerate(rows) if "vae_stem_kernel" in n]
        if len(stems) < 2:
            raise SystemExit("need at least one complete VAE chunk in the list")
        for n, us in rows[:stems[0]]:
            agg[n][0] += 1
            agg[n][1] += us
        complete = len(stems) - 1
        scale = n_chunks / complete
        for n, us in rows[stems[0]:stems[-1]]:
            agg[n][0] += scale
            agg[n][1] += us * scale
        note = (f"; {len(rows)} launches captured, {complete} complete VAE chunks averaged and scaled to {n_chunks} "
                f"(see the docstring of tools/summarize_launches.py)")
    else:
        for n, us in rows:
            agg[n][0] += 1
            agg[n][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"# {sum(v[0] for v in agg.values()):.0f} launches, {tot/1e3:.1f} ms of kernel time (cold-cache, serialised under "
          f"ncu: compare SHARES){note}")
    print(f"{'kernel':70s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8.0f} {us/1e3:10.2f} {100*us/tot:6.1f}%")
    gemm = sum(us for k, (n, us) in agg.items() if "gemm_" in k)
    print(f"# all gemm_* kernels: {100*gemm/tot:.1f} % of the step's kernel time")


if __name__ == "__main__":
    main()

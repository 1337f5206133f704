"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python profiles/summarize_launches.py gpurun_out/r01_launches_c3.csv > profiles/r01_launches_c3.txt
Kernels of this library are listed launch by launch; torch's data-generation kernels (outside the timed region) are
aggregated in one line."""
import collections
import csv
import sys


OURS = ("fused_small_kernel", "fused_tma_kernel", "impute_rows_kernel", "syrk_dmma_kernel", "reduce_", "loglike_kernel")


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3}
    ours, other = [], [0, 0.0]
    for r in rows[1:]:
        ms = float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
        name = r[ki].split("(")[0].replace("void ", "")
        if "boomgpu::" in name or any(k in name for k in OURS):
            ours.append((name, r[gi], r[bi], ms))
        else:
            other[0] += 1
            other[1] += ms
    tot = sum(o[3] for o in ours)
    print("# %s" % sys.argv[1])
    print("# launches of this library's kernels, in order (cold-cache, serialised by ncu: compare shares, not absolutes)")
    for name, g, b, ms in ours:
        print("%-50s grid=%-16s block=%-14s %10.4f ms" % (name, g, b, ms))
    agg = collections.OrderedDict()
    for name, _, _, ms in ours:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    print("\n# totals (this library): %.3f ms over %d launches" % (tot, len(ours)))
    for name, (c, ms) in agg.items():
        print("%-50s n=%-3d mean=%10.4f ms  share=%5.1f %%" % (name, c, ms / c, 100 * ms / tot))
    print("\n# other kernels in the process (torch: synthetic data generation, outside the timed region): %d launches, %.3f ms"
          % (other[0], other[1]))


if __name__ == "__main__":
    main()

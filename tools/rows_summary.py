"""profiles/rows_r02.md from the outputs of tools/rows_probe.py: the CUDA-event JSON (achieved GB/s against the
measured HBM copy peak) and the ncu metrics CSV of the same launches (DRAM bytes, duration).

    python tools/rows_summary.py gpurun_out/rows_r02_events.json gpurun_out/rows_r02_ncu.csv > profiles/rows_r02.md
"""
import collections
import csv
import json
import sys


def main():
    ev = json.load(open(sys.argv[1]))
    rows = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 10]
    h = rows[0]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault((r[ii], r[ki]), {})[r[mi]] = r[vi]
    peak = ev["hbm_peak_gbs"]
    print("# HBM-bound row / layout kernels at the headline shapes (round 2, B200)\n")
    print(f"`tools/rows_probe.py`: CUDA events (best of 5, 256 MB L2 flush between launches) against the measured HBM copy "
          f"peak of MEASURED_PEAKS.json ({peak:.0f} GB/s, read + write of a device copy), then the same launches under "
          "`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,... --clock-control none`.\n")
    print("## CUDA events\n")
    print("| kernel / shape | algorithmic bytes | ms | achieved GB/s | of measured HBM peak |\n|---|---:|---:|---:|---:|")
    for k in ev["kernels"]:
        print(f"| {k['kernel']} | {k['algorithmic_bytes'] / 1e6:.1f} MB | {k['ms']:.3f} | {k['achieved_gbs']:.0f} | "
              f"{k['frac_of_measured_hbm_peak']:.2f} |")
    print("\n## ncu (one launch each, cold caches, serialised)\n")
    print("| kernel | grid x block | duration us | dram read MB | dram write MB | dram GB/s | ncu dram % of peak |\n"
          "|---|---|---:|---:|---:|---:|---:|")
    f = lambda s: float(s.replace(",", ""))
    for (i, k), m in d.items():
        if "at::" in k or "distribution" in k:
            continue
        rd, wr, t = f(m["dram__bytes_read.sum"]), f(m["dram__bytes_write.sum"]), f(m["gpu__time_duration.sum"])
        name = k.split("(")[0].replace("void ", "")
        print(f"| `{name[:70]}` | {m['launch__grid_size']} x {m['launch__block_size']} | {t / 1e3:.1f} | {rd / 1e6:.1f} | "
              f"{wr / 1e6:.1f} | {(rd + wr) / t:.0f} | {m['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']} |")


if __name__ == "__main__":
    main()

"""Summarise gpurun_out/r2_ncu_ab_<tag>.csv files written by tools/ncu_ab.sh (development tool)."""
import csv
import sys
for tag in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f"gpurun_out/r2_ncu_ab_{tag}.csv")) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
    d = {}
    for r in rows[1:]:
        d.setdefault((int(r[ii]), r[ki].replace("void ", "")[:14]), {})[r[mi]] = float(r[vi].replace(",", ""))
    agg = {}
    for (i, k), v in d.items():
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += v["gpu__time_duration.sum"]; a[2] += v["dram__bytes_read.sum"]; a[3] += v["dram__bytes_write.sum"]
        a[4] += v["lts__t_sector_hit_rate.pct"]
    for k, a in agg.items():
        n = a[0]
        print(f"{tag:8s} {k:16s} n={n} time {a[1] / n / 1e3:7.2f} us  dram rd {a[2] / n / 1e6:7.1f} MB wr {a[3] / n / 1e6:6.1f} MB  "
              f"-> {(a[2] + a[3]) / a[1]:6.1f} GB/s  L2 hit {a[4] / n:5.1f} %")

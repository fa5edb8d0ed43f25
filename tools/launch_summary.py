"""profiles/*.md from an ncu launch list (--metrics gpu__time_duration.sum --csv):  python tools/launch_summary.py in.csv out.md "<command>" """
import collections
import csv
import sys

src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list -- `{cmd}`", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` on one B200 (per-launch times are cold-cache and",
       f"serialised: compare SHARES).  Raw list: `{src.replace('gpurun_out/', 'profiles/')}`.", "",
       "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[1] / tot * 100:.1f}% |")
pcg = {k: a[1] / a[0] for k, a in agg.items() if k.startswith(("void k_spmv<1", "void k_pcg_update", "void k_pcg_direction"))}
if len(pcg) == 3:
    s = sum(pcg.values())
    out += ["", "Inside one PCG iteration (the loop that is 99 % of a converged solve): " +
            ", ".join(f"`{k.replace('void ', '')}` {v:.1f} us = {v / s * 100:.1f} %" for k, v in pcg.items()) + "."]
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[5:24]))

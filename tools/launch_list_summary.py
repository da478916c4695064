"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel.
usage: python tools/launch_list_summary.py launches.csv "<command that was profiled>" > summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1.0)
    name = r[iK].split("(")[0]
    tot[name] += v
    cnt[name] += 1
print("ncu launch list (gpu__time_duration.sum, --clock-control none) of:", sys.argv[2] if len(sys.argv) > 2 else "?")
print("(per-launch times are cold-cache and serialised: compare SHARES, not absolutes; microbench kernels excluded from the step share)")
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'share_of_all':>12s}")
allt = sum(tot.values())
for k, v in tot.most_common():
    print(f"{k:60s} {cnt[k]:8d} {v:12.1f} {100 * v / allt:11.2f}%")
step = {k: v for k, v in tot.items() if "peak_kernel" not in k}
st = sum(step.values())
print("shares within the hot-path step (prep + table + map + lnBtSG + finalize):")
for k, v in sorted(step.items(), key=lambda kv: -kv[1]):
    print(f"  {k:58s} {100 * v / st:7.3f}%")

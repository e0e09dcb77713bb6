"""Condense an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) into per-kernel shares.

    python tools/launch_shares.py gpurun_out/launches_unet.csv "header line for the summary" > profiles/rNN_launch_shares.txt
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    header = sys.argv[2] if len(sys.argv) > 2 else path
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0]
        v, u = float(r[vi].replace(",", "")), r[ui]
        us = v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print(f"# {header}")
    print("# cmd: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python tools/ncu_cases.py unet")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:<62} n={a[0]:4d} total={a[1]:10.1f} us share={100 * a[1] / tot:5.1f}%")
    print(f"total {tot:.1f} us over {n} launches (cold-cache, serialised: compare shares)")


if __name__ == "__main__":
    main()

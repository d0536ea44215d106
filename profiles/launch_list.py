#!/usr/bin/env python
"""Turn an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`) into the
per-kernel table kept under profiles/.   usage: python profiles/launch_list.py X.csv "<cmd>" > profiles/<name>.txt"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        if r is hdr or r[mi] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ki]).strip()
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].replace("second", "s").replace("n s", "ns"), 1e-6)
        tot[name] += float(r[vi].replace(",", "")) * scale
        cnt[name] += 1
    total = sum(tot.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none: {sys.argv[2] if len(sys.argv) > 2 else ''} "
          "(per-launch times are cold-cache and serialised; shares are what counts)")
    print("%-70s %8s %10s %7s" % ("kernel", "launches", "total ms", "share"))
    for k, v in tot.most_common():
        print("%-70s %8d %10.3f %7.3f" % (k[:70], cnt[k], v, v / total))


if __name__ == "__main__":
    main()

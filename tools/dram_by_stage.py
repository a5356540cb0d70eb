#!/usr/bin/env python
"""Aggregates an ncu per-launch CSV (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control
none`, one profiled forward) by ResNet stage and layer role, using the plan's launch order (profile_forward.py --schedule-out).

  python tools/dram_by_stage.py name=gpurun_out/x.csv:gpurun_out/x_schedule.txt [name2=...]   ->  markdown on stdout"""
import collections
import csv
import gzip
import re
import sys


def load(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}
    for r in csv.DictReader(lines):
        d = per.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * mult.get(r["Metric Unit"], 1)
    return list(per.values())


def schedule(path):
    names = []
    for l in open(path):
        m = re.match(r"\s+issue (\w+) (\S+) (\d+) (\d+)", l)
        if m:
            names += ["conv1/s2d", "conv1"] if m.group(1) == "Conv1" else [m.group(2)]
    return names


def stage_of(name):
    m = re.match(r"res(\d)[a-z]\d*_branch(\w+)", name)
    if not m:
        return name.split("/")[0]
    b = m.group(2)
    return "res%s %s" % (m.group(1), "branch1" if b.startswith("1") else b[:2])


def main():
    tables = collections.OrderedDict()
    for spec in sys.argv[1:]:
        name, _, rest = spec.partition("=")
        c, _, s = rest.partition(":")
        rows, names = load(c), schedule(s)
        assert len(rows) == len(names), (name, len(rows), len(names))
        agg = collections.OrderedDict()
        for r, n in zip(rows, names):
            a = agg.setdefault(stage_of(n), [0.0, 0.0, 0.0, 0])
            a[0] += r["dram__bytes_read.sum"]
            a[1] += r["dram__bytes_write.sum"]
            a[2] += r["gpu__time_duration.sum"]
            a[3] += 1
        tables[name] = agg
    keys = list(next(iter(tables.values())).keys())
    print("| layers | " + " | ".join("%s: launches / DRAM read+write MB / kernel time us" % n for n in tables) + " |")
    print("|---|" + "---|" * len(tables))
    for k in keys:
        print("| %s | " % k + " | ".join("%d / %.0f + %.0f / %.0f" % (t[k][3], t[k][0] / 1e6, t[k][1] / 1e6, t[k][2]) for t in tables.values()) + " |")
    print("| **forward** | " + " | ".join("%d / %.2f + %.2f GB = **%.1f GB** / %.0f" % (sum(a[3] for a in t.values()), sum(a[0] for a in t.values()) / 1e9,
                                                                                   sum(a[1] for a in t.values()) / 1e9, sum(a[0] + a[1] for a in t.values()) / 1e9,
                                                                                   sum(a[2] for a in t.values())) for t in tables.values()) + " |")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""ncu launch-list CSV (tools/gpu_launchlist.sh, --graph-profiling node) -> markdown table of ONE forward pass + shares by kernel family.
usage: launchlist_md.py launches.csv "title" [first-kernel-substring]   (a forward = rows between two launches of the first kernel)"""
import collections, csv, re, sys


def load(fn):
    with open(fn) as f:
        lines = [l for l in f if not l.startswith("==")]
    cur = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = cur.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r.get("Grid Size", "")})
        d[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    return list(cur.values())


def us(v):
    x, u = v
    return x / 1000 if u == "ns" else x * 1000 if u == "ms" else x


def mb(v):
    x, u = v
    return x / 1e6 if u == "byte" else x / 1e3 if u == "Kbyte" else x * 1e3 if u == "Gbyte" else x


def main():
    rows = load(sys.argv[1])
    title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    first = sys.argv[3] if len(sys.argv) > 3 else "pack_smallc"
    idx = [i for i, r in enumerate(rows) if first in r["name"]]
    if len(idx) < 2:
        raise SystemExit("need two launches of '%s' to delimit a forward" % first)
    # the input's abs-max kernel (fp32-parity mode) precedes the first pack
    a, b = idx[0], idx[1]
    if a > 0 and "absmax" in rows[a - 1]["name"]:
        a, b = a - 1, b - 1
    fwd = rows[a:b]
    tot = sum(us(r["gpu__time_duration.sum"]) for r in fwd)
    print("# %s\n" % title)
    print("%d kernels, sum %.1f us (cold-cache, serialised: compare SHARES, not absolutes).\n" % (len(fwd), tot))
    print("| kernel | grid | us | share | DRAM read MB | DRAM write MB | tensor pipe active % |\n|---|---|---|---|---|---|---|")
    fam = collections.Counter()
    for r in fwd:
        name = re.sub(r"\(.*", "", r["name"]).replace("void ", "").replace("b200::", "")
        d = us(r["gpu__time_duration.sum"])
        fam[re.sub(r"<.*", "", name)] += d
        print("| %s | %s | %.1f | %.1f%% | %.1f | %.1f | %.1f |" % (name, r["grid"], d, 100 * d / tot, mb(r["dram__bytes_read.sum"]), mb(r["dram__bytes_write.sum"]),
                                                                  r["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]))
    print("| **sum** | | %.1f | | | | |\n" % tot)
    print("Shares by kernel: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in fam.most_common()))


if __name__ == "__main__":
    main()

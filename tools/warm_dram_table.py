"""Per-kernel table of a WARM forward from `ncu --cache-control none --graph-profiling node` CSV logs
(metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum): the last complete forward
(51 launches from the prep kernel) of each file, side by side.  usage: warm_dram_table.py a.csv [b.csv ...]"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rd:
        if len(r) < len(hdr):
            continue
        d = per.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u, m = r[ix["Metric Unit"]], r[ix["Metric Name"]]
        if "bytes" in m:
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        if "time" in m:
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}[u]
        d[m] = v
    return list(per.values())


def short(n):
    m = re.search(r"(\w+)<([^>]*)>", n)
    return (m.group(1)[:12] + "<" + m.group(2).replace(" ", "") + ">") if m else n[:30]


def last_forward(launches, n=51):
    idx = [i for i, d in enumerate(launches) if "prep" in d["name"]]
    for i in reversed(idx):
        if i + n <= len(launches):
            return launches[i:i + n]
    raise SystemExit("no complete forward in the capture")


def main():
    fwd = [last_forward(load(p)) for p in sys.argv[1:]]
    print("kernel".ljust(46) + " | ".join("   us   rd MB   wr MB" for _ in fwd))
    tot = [[0.0, 0.0, 0.0] for _ in fwd]
    for row in zip(*fwd):
        cells = []
        for k, x in enumerate(row):
            t, r, w = x["gpu__time_duration.sum"], x["dram__bytes_read.sum"] / 1e6, x["dram__bytes_write.sum"] / 1e6
            tot[k][0] += t; tot[k][1] += r; tot[k][2] += w
            cells.append(f"{t:6.1f} {r:7.1f} {w:7.1f}")
        print(short(row[0]["name"]).ljust(46) + " | ".join(cells))
    print("sum".ljust(46) + " | ".join(f"{t:6.1f} {r:7.1f} {w:7.1f}" for t, r, w in tot))


if __name__ == "__main__":
    main()

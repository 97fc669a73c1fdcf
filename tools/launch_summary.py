"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py FILE [last_n_launches_per_kernel]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, agg, order = None, collections.OrderedDict(), []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v *= {"nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3}.get(u, 1e-6)
        agg.setdefault(name, []).append(v)
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    tot = 0.0
    out = []
    for k, v in agg.items():
        if steps:
            per = len(v) // steps if len(v) >= steps else len(v)
            v = v[-per:]                      # the last step's launches
        out.append((sum(v), len(v), k))
        tot += sum(v)
    for t, n, k in sorted(out, reverse=True):
        print("%-48s %9.3f ms %5d  %5.1f%%" % (k[:48], t, n, 100 * t / tot))
    print("total %.3f ms" % tot)


if __name__ == "__main__":
    main()

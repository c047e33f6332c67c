"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and grid."""
import collections, csv, sys
def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        agg.setdefault((r["Kernel Name"][:60], r["Grid Size"]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("launches %d, total %.1f us (cold-cache, serialised: compare shares)" % (sum(len(v) for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-62s grid %-16s n=%4d avg %9.1f us total %10.1f us %5.1f%%" % (k[0], k[1], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
if __name__ == "__main__":
    main(sys.argv[1])

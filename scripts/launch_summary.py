"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / mean."""
import csv, sys, collections
def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    agg = collections.OrderedDict()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ms
    print("| kernel | launches | total ms | ms / launch |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[:120]} | {n} | {t:.2f} | {t / n:.3f} |")
if __name__ == "__main__":
    main(sys.argv[1])

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import collections, csv, re, sys

def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg[name][0] += 1; agg[name][1] += v; tot += v
    print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches (cold-cache, serialised: compare SHARES)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.1f} us {100 * t / tot:5.1f}%  x{n:4d}  {k[:100]}")

if __name__ == "__main__":
    main(sys.argv[1])

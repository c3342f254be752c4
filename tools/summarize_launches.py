"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections, csv, re, sys

def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"total {T/1e3:.3f} ms over {sum(cnt.values())} launches")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v/1e3:9.3f} ms {100*v/T:5.1f}%  n={cnt[k]:5d}  avg={v/cnt[k]:8.1f} us  {k[:100]}")

main(sys.argv[1])

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table."""
import collections
import csv
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(row["Metric Unit"], 1.0)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    rows = [f"# {path}: {sum(cnt.values())} launches, {T / 1e3:.2f} ms of kernel time (ncu: cold cache, serialised — compare shares)",
            f"{'kernel':34s} {'launches':>8s} {'total_us':>11s} {'avg_us':>9s} {'share':>6s}"]
    for k, v in tot.most_common():
        rows.append(f"{k[:34]:34s} {cnt[k]:8d} {v:11.1f} {v / cnt[k]:9.2f} {v / T:6.3f}")
    text = "\n".join(rows) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(*sys.argv[1:3])

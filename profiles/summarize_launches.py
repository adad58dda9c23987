"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel:
launches, total device time, share of the captured window.  Usage:
    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.md
(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("mmt::", "")
        name = re.sub(r"<.*", lambda m: m.group(0) if len(m.group(0)) < 24 else "<...>", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "nsecond": 1.0, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
        rows.append((name, ns, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, set()])
        a[0] += 1
        a[1] += ns
        if len(a[2]) < 4:
            a[2].add(f"{grid}x{block}")
    total = sum(a[1] for a in agg.values())
    print(f"source: {path}; {len(rows)} launches, {total / 1e6:.3f} ms device time in the captured window\n")
    print("| kernel | launches | total ms | share | mean us | grid x block (examples) |")
    print("|---|---:|---:|---:|---:|---|")
    for name, (n, ns, shapes) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.2f} | {'; '.join(sorted(shapes))} |")


if __name__ == "__main__":
    main(sys.argv[1])

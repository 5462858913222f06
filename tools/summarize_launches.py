"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import io
import re
import sys


def main(path, iters):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"])[:64]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v *= {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("s3g::") or "s3g::" in k)
    print(f"{path}: {len(rows)} launches, {tot / 1e6 / iters:.3f} ms/iter total, s3g kernels {ours / 1e6 / iters:.3f} ms/iter")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        print(f"  {t / 1e6 / iters:8.3f} ms/iter  n={n:3d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of device time).
Usage: python tools/summarize_launches.py gpurun_out/launches.csv [first] [last]  > profiles/<name>.md"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(?:zsg::)?(?:\(anonymous namespace\)::)?([A-Za-z0-9_:]+)(<[^(]*>)?\(", name)
    if name.startswith("at::") or "at::native" in name:
        m2 = re.search(r"(FillFunctor|CopyFunctor|direct_copy|index|cat|elementwise)[A-Za-z_]*", name)
        return "torch::" + (m2.group(0) if m2 else name[:40])
    if m:
        t = m.group(2) or ""
        return m.group(1).replace("zsg::", "") + (t if len(t) < 24 else "")
    return name[:60]


def main():
    path = sys.argv[1]
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    last = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        i = int(r["ID"])
        if first <= i < last:
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))
    tot = sum(t for _, t, _, _ in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, t, _, _ in rows:
        agg[k][0] += 1
        agg[k][1] += t
    print(f"launches {len(rows)} (IDs {first}..{min(last, first + len(rows))}), total device time {tot / 1e3:.3f} ms "
          f"(ncu per-launch times: cold cache, serialised)\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.1f} |")


if __name__ == "__main__":
    main()

"""Markdown summary + traffic JSON of an ncu launch list of ONE step (tools/profile_step.py under
`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`).
Usage: python tools/summarize_step.py <csv> <tag e.g. fp32@bs64> <out.md> [traffic.json to update]"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

path, tag, out_md = sys.argv[1:4]
traffic_json = sys.argv[4] if len(sys.argv) > 4 else None
lines = [ln for ln in open(path) if ln.startswith('"')]
per = defaultdict(dict)
names = {}
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3,
             "msecond": 1e6}.get(r["Metric Unit"], 1)
    per[r["ID"]][r["Metric Name"]] = v * scale
    names[r["ID"]] = r["Kernel Name"]


def short(name):
    name = re.sub(r"^void ", "", name)
    if "at::native" in name or name.startswith("at::"):
        m = re.search(r"(FillFunctor|CopyFunctor|direct_copy|index|cat|elementwise|reduce)[A-Za-z_]*", name)
        return "torch::" + (m.group(0) if m else name[:40])
    m = re.match(r"(?:zsg::)?([A-Za-z0-9_]+)(<[^(]*>)?\(", name)
    if m:
        t = m.group(2) or ""
        return m.group(1) + (t if len(t) < 24 else "")
    return name[:60]


agg = defaultdict(lambda: [0, 0.0, 0.0])
for i, mets in per.items():
    a = agg[short(names[i])]
    a[0] += 1
    a[1] += mets.get("gpu__time_duration.sum", 0.0)
    a[2] += mets.get("dram__bytes_read.sum", 0.0) + mets.get("dram__bytes_write.sum", 0.0)
tot_ns = sum(a[1] for a in agg.values())
tot_b = sum(a[2] for a in agg.values())
with open(out_md, "w") as f:
    f.write(f"# ncu launch list of ONE training step, {tag}\n\n"
            f"`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
            f"--clock-control none --csv python tools/profile_step.py {tag.replace('@bs', ' ')}` (eager launches; per-launch times are "
            "cold-cache and serialised: compare SHARES with bench.py's CUDA-event numbers, not absolutes)\n\n"
            f"launches {sum(a[0] for a in agg.values())}, device time {tot_ns / 1e6:.2f} ms, DRAM traffic {tot_b / 1e9:.1f} GB\n\n"
            "| kernel | launches | total us | share | avg us | DRAM MB / launch | GB/s |\n|---|---|---|---|---|---|---|\n")
    for k, (n, ns, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {ns / 1e3:.1f} | {100 * ns / tot_ns:.1f}% | {ns / n / 1e3:.1f} | {b / n / 1e6:.1f} | {b / max(ns, 1):.0f} |\n")
if traffic_json:
    d = json.load(open(traffic_json)) if os.path.exists(traffic_json) else {}
    suffix = "@" + tag
    for key in [k for k in d if k.endswith(suffix)]:     # a re-run replaces this tag's entries
        del d[key]
    for k, (n, ns, b) in agg.items():
        base = re.sub(r"<.*", "", k)
        m = re.match(r"conv_tc_async_kernel<\d+, (\w+)", k)
        if m and m.group(1) in ("1", "true"):              # second template argument: the bf16 operand path
            base = "conv_bf16_kernel"
        key = f"{base}{suffix}"
        e = d.setdefault(key, dict(launches=0, dram_bytes=0.0, ns=0.0))
        e["launches"] += n
        e["dram_bytes"] += b
        e["ns"] += ns
    for e in d.values():
        if isinstance(e, dict) and e.get("launches"):
            e["dram_bytes_per_launch"] = e["dram_bytes"] / e["launches"]
            e["us_per_launch"] = e["ns"] / e["launches"] / 1e3
    d["_source"] = "tools/summarize_step.py over ncu launch lists of tools/profile_step.py (one step each)"
    json.dump(d, open(traffic_json, "w"), indent=1, sort_keys=True)
print(open(out_md).read())

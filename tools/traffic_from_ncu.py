"""profiles/r01_traffic.json from an ncu CSV with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum:
per kernel (name as used by bench.py's launch profiler) the DRAM bytes per launch averaged over the captured launches."""
import csv, json, re, sys
from collections import defaultdict
path, out = sys.argv[1], sys.argv[2]
lines = [ln for ln in open(path) if ln.startswith('"')]
per = defaultdict(lambda: defaultdict(float))
for r in csv.DictReader(lines):
    m = re.search(r"(conv_tc_async_kernel|wgrad_tc_async_kernel|conv_tc_kernel|wgrad_tc_kernel|conv_tc_generic_kernel|[a-z_0-9]+_kernel)", r["Kernel Name"])
    if not m:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
    per[(m.group(1), r["ID"])][r["Metric Name"]] = v * scale
agg = defaultdict(lambda: dict(launches=0, dram_bytes=0.0, ns=0.0))
for (k, _), mets in per.items():
    a = agg[k]
    a["launches"] += 1
    a["dram_bytes"] += mets.get("dram__bytes_read.sum", 0.0) + mets.get("dram__bytes_write.sum", 0.0)
    a["ns"] += mets.get("gpu__time_duration.sum", 0.0)
res = {k: dict(launches=a["launches"], dram_bytes_per_launch=a["dram_bytes"] / a["launches"], us_per_launch=a["ns"] / a["launches"] / 1e3,
               dram_gbs=a["dram_bytes"] / max(a["ns"], 1.0)) for k, a in agg.items()}
res["_source"] = "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (python bench.py --steps 1 --warmup 3)"
json.dump(res, open(out, "w"), indent=1, sort_keys=True)
for k, a in sorted(res.items(), key=lambda kv: -(kv[1]["dram_bytes_per_launch"] * kv[1]["launches"]) if isinstance(kv[1], dict) else 0):
    if isinstance(a, dict):
        print(f"{k:32s} launches {a['launches']:4d}  dram/launch {a['dram_bytes_per_launch']/1e6:9.2f} MB  {a['us_per_launch']:8.1f} us  {a['dram_gbs']:7.1f} GB/s")

"""Summarise gpurun_out/b3d_*.json and the 3D launch list."""
import collections
import csv
import json
import sys

for f in ("gpurun_out/b3d_16.json", "gpurun_out/b3d_32.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms %.1f e2e %.0f k %.1f launches %d frac %.3f" % (
            j["value"], j["ms_per_step"], j["e2e"]["value"], j["config"]["mean_pcg_iterations"],
            j["gpu_launches"], j["roofline"]["frac"]))
    except Exception as e:
        print(f, "ERR", e)
name = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches_3d.csv"
with open(name) as f:
    lines = [l for l in f if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] in ("ns", "nsecond") else v
    tot[k] += v
    cnt[k] += 1
T = sum(tot.values())
for k, v in tot.most_common(12):
    print("%-40s n=%4d total %9.1f us  avg %8.1f us  %5.1f%%" % (k[:40], cnt[k], v, v / cnt[k], 100 * v / T))

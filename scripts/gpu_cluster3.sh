#!/bin/bash
# cluster / DSMEM tier: A/B of the two kernel flavours + ncu --set full captures with source for the line profile
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 300 -x -k "cluster" > gpurun_out/pytest_cluster.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_cluster.log
tail -4 gpurun_out/pytest_cluster.log
for wv in "cfg1 0" "cfg1 4" "cfg1x64 0" "cfg1x64 4"; do
  set -- $wv
  timeout 200 python bench.py --workload $1 --variant $2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_$1_v$2.json 2> gpurun_out/ab_$1_v$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_$1_v$2.json"))
    print("$1 v$2", "solves/s %.0f"%d["value"], "k %.1f"%d["config"]["mean_pcg_iterations"], "solve_ms %.3f"%d["roofline"]["kernel_ms_per_launch"], "step_ms %.3f"%d["ms_per_step"], "launches", d["gpu_launches"], "frac %.3f"%d["roofline"]["frac"])
except Exception as e:
    print("$1 v$2 FAILED", e); print(open("gpurun_out/ab_$1_v$2.err").read()[-800:])
PY
done
for v in 4 0; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_cluster -c 1 -f -o gpurun_out/cluster_cfg1x64_v$v python bench.py --workload cfg1x64 --cells 1024 --variant $v --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_cluster_v$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep

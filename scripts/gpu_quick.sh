#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for wv in "target 0" "target 5"; do
  set -- $wv
  timeout 300 python bench.py --workload $1 --cells 5920 --variant $2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sweep_$1_v$2.json 2> gpurun_out/sweep_$1_v$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_$1_v$2.json"))
    print("$1 v$2", "solves/s %.0f"%d["value"], "k %.1f"%d["config"]["mean_pcg_iterations"], "solve_ms %.2f"%d["roofline"]["kernel_ms_per_launch"], "step_ms %.2f"%d["ms_per_step"])
except Exception as e:
    print("$1 v$2 FAILED", e); print(open("gpurun_out/sweep_$1_v$2.err").read()[-600:])
PY
done
timeout 900 python bench.py > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err
cut -c1-300 gpurun_out/bench_target.json
tail -3 gpurun_out/bench_target.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 9 --csv --log-file gpurun_out/launches.csv python bench.py --workload target --cells 5920 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1; grep -o "\"[a-z_]*kernel[^\"]*\".*" gpurun_out/launches.csv | cut -d, -f1,11- | tail -4

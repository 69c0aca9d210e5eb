#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for w in target cfg4 cfg2 cfg1; do
  timeout 300 python bench.py --workload $w --cells 5920 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sweep_${w}.json 2> gpurun_out/sweep_${w}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_${w}.json"))
    print("$w", "solves/s %.0f"%d["value"], "k %.1f"%d["config"]["mean_pcg_iterations"], "solve_ms %.2f"%d["roofline"]["kernel_ms_per_launch"], "step_ms %.2f"%d["ms_per_step"])
except Exception as e:
    print("$w FAILED", e); print(open("gpurun_out/sweep_${w}.err").read()[-800:])
PY
done
timeout 900 python bench.py > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err
cat gpurun_out/bench_target.json | cut -c1-300
tail -3 gpurun_out/bench_target.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches.csv python bench.py --workload target --cells 5920 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
grep -o '"[a-z_]*kernel[^"]*".*' gpurun_out/launches.csv | cut -d, -f1,11- | tail -6

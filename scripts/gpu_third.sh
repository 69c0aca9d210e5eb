#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
for w in target cfg4 cfg2; do
 for v in 0 1; do
  timeout 300 python bench.py --workload $w --cells 5920 --variant $v --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sweep_${w}_v$v.json 2> gpurun_out/sweep_${w}_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_${w}_v$v.json"))
    print("$w v$v", "solves/s %.0f"%d["value"], "k %.1f"%d["config"]["mean_pcg_iterations"], "solve_ms %.2f"%d["roofline"]["kernel_ms_per_launch"], "step_ms %.2f"%d["ms_per_step"])
except Exception as e:
    print("$w v$v FAILED", e); print(open("gpurun_out/sweep_${w}_v$v.err").read()[-800:])
PY
 done
done
timeout 900 python bench.py > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err
cat gpurun_out/bench_target.json | cut -c1-400
tail -3 gpurun_out/bench_target.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_bpx -c 1 -o gpurun_out/prof_solve_bpx2 python bench.py --workload target --cells 1184 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls gpurun_out | head -40

#!/bin/bash
# final binary of the round: default bench + reference arm, n = 32 kernel evidence, variant 7 A/B at n = 64
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err; cut -c1-170 gpurun_out/bench_target.json; tail -2 gpurun_out/bench_target.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-150 gpurun_out/bench_reference.json
for v in 0 7; do
  timeout 200 python bench.py --workload target --cells 5920 --variant $v --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/v7_$v.json 2> gpurun_out/v7_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/v7_$v.json")); print("target 5920 cells variant $v", "solves/s %.0f"%d["value"], "k %.2f"%d["config"]["mean_pcg_iterations"], "solve_ms %.3f"%d["roofline"]["kernel_ms_per_launch"])
PY
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --workload cfg4 --cells 11840 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_cfg4.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:solve_bpx_kernel -c 1 -f -o gpurun_out/bpx5_cfg4_ldl python bench.py --workload cfg4 --cells 2368 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_bpx5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2

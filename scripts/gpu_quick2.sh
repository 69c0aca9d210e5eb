#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "cluster or default_run_all_64 or survey_crosscheck or general_quadrilateral or fused_and_unfused or streamed_tier_256 or host_driver" > gpurun_out/pytest_sub.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sub.log
tail -4 gpurun_out/pytest_sub.log
timeout 300 python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; cut -c1-170 gpurun_out/bench_cfg1.json; tail -2 gpurun_out/bench_cfg1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg1.csv python bench.py --workload cfg1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_cfg1.log 2>&1
grep -c "element_matrix" gpurun_out/launches_cfg1.csv

#!/bin/bash
# First GPU session: correctness, smoke, a first bench line, variant sweep, ncu launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
for v in 0 1 2 3; do
  timeout 300 python bench.py --workload target --cells 5920 --variant $v --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sweep_target_v$v.json 2> gpurun_out/sweep_target_v$v.err
done
for v in 0 1 2; do
  timeout 300 python bench.py --workload cfg4 --cells 11840 --variant $v --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/sweep_cfg4_v$v.json 2> gpurun_out/sweep_cfg4_v$v.err
done
grep -h -o '"value": [0-9.e+]*\|"variant": [0-9]\|"kernel_ms_per_launch": [0-9.]*\|"mean_pcg_iterations": [0-9.]*' gpurun_out/sweep_*.json | paste - - - - 
timeout 900 python bench.py > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err
cat gpurun_out/bench_target.json | cut -c1-1500
tail -3 gpurun_out/bench_target.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --workload target --cells 5920 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_smem -c 1 -o gpurun_out/prof_solve_v0 python bench.py --workload target --cells 1184 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out

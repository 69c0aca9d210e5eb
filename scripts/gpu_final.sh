#!/bin/bash
# round-end validation: full GPU test suite, A/B of the n = 128 kernels, bench lines, ncu evidence
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for wv in "cfg1 0" "cfg1 4" "cfg1 2" "cfg1x64 0" "cfg1x64 4"; do
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
timeout 600 python bench.py > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err; cut -c1-200 gpurun_out/bench_target.json; tail -2 gpurun_out/bench_target.err
timeout 300 python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; cut -c1-200 gpurun_out/bench_cfg1.json; tail -2 gpurun_out/bench_cfg1.err
timeout 300 python bench.py --workload cfg1x64 --steps 3 > gpurun_out/bench_cfg1x64.json 2> gpurun_out/bench_cfg1x64.err; cut -c1-200 gpurun_out/bench_cfg1x64.json; tail -2 gpurun_out/bench_cfg1x64.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_cfg1.csv python bench.py --workload cfg1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_cfg1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_cluster -c 1 -f -o gpurun_out/cluster_cfg1x64_final python bench.py --workload cfg1x64 --cells 1024 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_cluster_final.log 2>&1
MSB_LIBRARY=$PWD/mpi_parallel_multiscale_diffusion_fem_b200/libmsfem_basis_prof.so timeout 120 python scripts/stage_timers.py cfg1x64 1024 2>&1 | tee gpurun_out/stage_timers_cfg1x64.txt
timeout 150 compute-sanitizer --tool memcheck python scripts/sanitize_case.py 7 > gpurun_out/sanitizer_memcheck_cluster.txt 2>&1; tail -3 gpurun_out/sanitizer_memcheck_cluster.txt
timeout 150 compute-sanitizer --tool racecheck python scripts/sanitize_case.py 7 > gpurun_out/sanitizer_racecheck_cluster.txt 2>&1; tail -3 gpurun_out/sanitizer_racecheck_cluster.txt
ls -la gpurun_out/*.ncu-rep

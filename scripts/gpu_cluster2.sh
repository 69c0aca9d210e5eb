#!/bin/bash
# cluster / DSMEM tier, quick loop: parity tests, A/B against the HBM-streamed kernels, stage timers, sanitizer
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 300 -x \
  -k "cluster or default_run_all_64 or survey_crosscheck or general_quadrilateral or fused_and_unfused" \
  > gpurun_out/pytest_cluster.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_cluster.log
tail -5 gpurun_out/pytest_cluster.log
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
MSB_LIBRARY=$PWD/mpi_parallel_multiscale_diffusion_fem_b200/libmsfem_basis_prof.so timeout 120 python scripts/stage_timers.py cfg1x64 1024 2>&1 | tee gpurun_out/stage_timers_cfg1x64.txt
MSB_LIBRARY=$PWD/mpi_parallel_multiscale_diffusion_fem_b200/libmsfem_basis_prof.so timeout 120 python scripts/stage_timers.py cfg1x64 1024 4 2>&1 | tee gpurun_out/stage_timers_cfg1x64_v4.txt
timeout 150 compute-sanitizer --tool memcheck python scripts/sanitize_case.py 7 > gpurun_out/sanitizer_memcheck_cluster.txt 2>&1; tail -2 gpurun_out/sanitizer_memcheck_cluster.txt
timeout 150 compute-sanitizer --tool racecheck python scripts/sanitize_case.py 7 > gpurun_out/sanitizer_racecheck_cluster.txt 2>&1; tail -2 gpurun_out/sanitizer_racecheck_cluster.txt

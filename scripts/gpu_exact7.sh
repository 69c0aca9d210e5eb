#!/bin/bash
# banded LDL^T inverse of the 7x7 level (n = 32 kernel, n = 64 variant 7): parity + timing
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "kernel_variants_agree or full_size_invariants or bases_match_oracle or tensor_memory_kernel or streamed_tier_equals or no_convergence or set_global_weights or handle_reuse" > gpurun_out/pytest_exact7.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_exact7.log
tail -4 gpurun_out/pytest_exact7.log
for wv in "cfg4 0" "cfg2 0" "cfg4 3"; do
  set -- $wv
  timeout 200 python bench.py --workload $1 --variant $2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/x7_$1_v$2.json 2> gpurun_out/x7_$1_v$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/x7_$1_v$2.json"))
    print("$1 v$2", "solves/s %.0f"%d["value"], "k %.1f"%d["config"]["mean_pcg_iterations"], "solve_ms %.3f"%d["roofline"]["kernel_ms_per_launch"], "step_ms %.3f"%d["ms_per_step"])
except Exception as e:
    print("$1 v$2 FAILED", e); print(open("gpurun_out/x7_$1_v$2.err").read()[-800:])
PY
done
MSB_LIBRARY=$PWD/mpi_parallel_multiscale_diffusion_fem_b200/libmsfem_basis_prof.so python scripts/stage_timers.py cfg4 2368 0 2>&1 | tee gpurun_out/stage_timers_cfg4_ldl.txt
timeout 150 compute-sanitizer --tool racecheck python scripts/sanitize_case.py 5 > gpurun_out/sanitizer_racecheck_n32.txt 2>&1; tail -3 gpurun_out/sanitizer_racecheck_n32.txt
timeout 150 compute-sanitizer --tool memcheck python scripts/sanitize_case.py 5 > gpurun_out/sanitizer_memcheck_n32.txt 2>&1; tail -2 gpurun_out/sanitizer_memcheck_n32.txt

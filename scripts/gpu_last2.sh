#!/bin/bash
# ncu evidence of the round-end default kernel at n = 64 (solve_bpx_tm_kernel<512,true>) + its stage timers
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_target5920.csv python bench.py --workload target --cells 5920 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_target.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:solve_bpx_tm -c 1 -f -o gpurun_out/bpx_tm_exact_1184 python bench.py --workload target --cells 1184 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_target.log 2>&1
MSB_LIBRARY=$PWD/mpi_parallel_multiscale_diffusion_fem_b200/libmsfem_basis_prof.so timeout 120 python scripts/stage_timers.py target 1184 2>&1 | tee gpurun_out/stage_timers_target_exact.txt
ls -la gpurun_out/*.ncu-rep | tail -1

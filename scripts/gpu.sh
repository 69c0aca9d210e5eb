#!/bin/bash
# gpu.sh -- the command sequences sent to the GPU box (gpurun -- 'bash scripts/gpu.sh STEP [STEP...]').
# Every step writes under gpurun_out/; what is kept as evidence is copied to profiles/ afterwards.
#   tests            python -m pytest tests -m gpu
#   tests-new        only the tests added in round 2 (fast)
#   bench            python bench.py (defaults: target + per_config + e2e legs + CPU baseline)
#   bench-quick      bench.py target only, no per_config / CPU baseline
#   ref              python bench.py --impl reference
#   probes           scripts/probes/onchip_peaks (FP64 FMA, shared-memory crossbar, HBM copy)
#   ncu-launches     ncu launch list of a 5920-cell target step
#   ncu-full         ncu --set full of the target solve kernel on 1184 cells
#   timers           clock64 stage timers of the target solve kernel (profiling build)
#   ab:V1,V2,..      bench.py --variant V for each V on 5920 target cells (A/B of kernel variants)
#   wl:NAME[:VAR]    bench.py --workload NAME [--variant VAR]
#   abw:NAME:CELLS:V1,V2   A/B of variants on a slice of any workload
#   ncuwl:NAME:CELLS:REGEX[:SKIP]   ncu --set full of one launch of a kernel of another workload
#   launches:NAME:CELLS    ncu launch list of one step of a workload slice
#   (NCU_CELLS=5920 ncu-full captures a launch larger than the L2, so that dram__bytes shows the real traffic)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PY=python
for step in "$@"; do
  echo "=== $step ==="
  case "$step" in
    tests)
      timeout 2400 $PY -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_tail.txt ;;
    tests-new)
      timeout 1200 $PY -m pytest tests -m gpu -x -q -k "l9 or l6 or cfg5 or bulk or device_results or invalidates or unsymmetric or independent or fused or table" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_new_tail.txt ;;
    bench)
      timeout 1500 $PY bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 6000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err ;;
    bench-quick)
      timeout 600 $PY bench.py --no-per-config --no-cpu-baseline --no-bases > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err ;;
    ref)
      timeout 900 $PY bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json ;;
    probes)
      (cd scripts/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o onchip_peaks onchip_peaks.cu) && timeout 300 scripts/probes/onchip_peaks | tee gpurun_out/onchip_peaks.json ;;
    latency)
      (cd scripts/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_probe latency_probe.cu) && timeout 120 scripts/probes/latency_probe | tee gpurun_out/latency_probe.json ;;
    ncu-launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_target5920.csv $PY bench.py --workload target --cells 5920 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_target.log 2>&1; tail -8 gpurun_out/launches_target5920.csv ;;
    ncu-full)
      V=${NCU_VARIANT:-0}
      timeout 900 ncu --set full --import-source on --clock-control none -k regex:'solve_fused|solve_bpx_tm' -c 1 -f -o gpurun_out/solve_n64_v${V}_${NCU_CELLS:-1184} $PY bench.py --workload target --cells ${NCU_CELLS:-1184} --steps 1 --warmup 3 --variant $V --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_target.log 2>&1
      ncu -i gpurun_out/solve_n64_v${V}_${NCU_CELLS:-1184}.ncu-rep --page raw --csv > gpurun_out/ncu_full_solve_n64_v${V}_${NCU_CELLS:-1184}.csv 2>/dev/null; ls -la gpurun_out/*.ncu-rep | tail -2 ;;
    timers)
      MSB_LIBRARY=$PWD/mpi_parallel_multiscale_diffusion_fem_b200/libmsfem_basis_prof.so timeout 300 $PY scripts/stage_timers.py target 1184 ${TIMER_VARIANT:-0} 2>&1 | tee gpurun_out/stage_timers_target_1184_v${TIMER_VARIANT:-0}.txt ;;
    ab:*)
      for v in $(echo "${step#ab:}" | tr ',' ' '); do
        timeout 90 $PY bench.py --workload target --cells 5920 --steps 5 --warmup 3 --variant $v --no-cpu-baseline --no-e2e > gpurun_out/ab_target5920_v$v.json 2> gpurun_out/ab_v$v.err
        $PY -c "import json,sys; d=json.load(open('gpurun_out/ab_target5920_v$v.json')); print('variant $v: %.1f k solves/s, %.3f ms/step, solve kernel %.3f ms, k=%.2f' % (d['value']/1e3, d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['config']['mean_pcg_iterations']))" || tail -3 gpurun_out/ab_v$v.err
      done ;;
    abt:*)
      # A/B through the streamed / cluster tier: abt:V1,V2
      for v in $(echo "${step#abt:}" | tr ',' ' '); do
        timeout 90 $PY bench.py --workload target --cells 5920 --steps 5 --warmup 3 --variant $v --tier 2 --no-cpu-baseline --no-e2e > gpurun_out/abt_target5920_v$v.json 2> gpurun_out/abt_v$v.err
        $PY -c "import json,sys; d=json.load(open('gpurun_out/abt_target5920_v$v.json')); print('tier 2 variant $v: %.1f k solves/s, %.3f ms/step, solve kernel %.3f ms, k=%.2f' % (d['value']/1e3, d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['config']['mean_pcg_iterations']))" || tail -3 gpurun_out/abt_v$v.err
      done ;;
    wl:*)
      spec="${step#wl:}"; name="${spec%%:*}"; var=0; [[ "$spec" == *:* ]] && var="${spec#*:}"
      timeout 900 $PY bench.py --workload $name --variant $var --no-cpu-baseline --no-bases > gpurun_out/bench_${name}_v$var.json 2> gpurun_out/bench_${name}_v$var.err
      $PY -c "import json; d=json.load(open('gpurun_out/bench_${name}_v$var.json')); print('$name v$var: %.1f k solves/s, %.3f ms/step, kernel %.3f ms, k=%.2f, model frac %.3f, clocks %s' % (d['value']/1e3, d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['config']['mean_pcg_iterations'], d['roofline']['frac'], d['clocks']))" || tail -3 gpurun_out/bench_${name}_v$var.err ;;
    ncuwl:*)
      # ncuwl:WORKLOAD:CELLS:KERNEL_REGEX[:SKIP] -- ncu --set full of one launch of a kernel of another workload
      IFS=: read -r _ name cells rx skip <<< "$step"
      timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$rx" -s ${skip:-20} -c 1 -f -o gpurun_out/ncu_${name}_${rx} $PY bench.py --workload $name --cells $cells --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_${name}.log 2>&1
      ncu -i gpurun_out/ncu_${name}_${rx}.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_${rx}.csv 2>/dev/null; ls -la gpurun_out/ncu_${name}_${rx}.* ;;
    launches:*)
      # launches:WORKLOAD:CELLS -- ncu launch list (durations) of one step
      IFS=: read -r _ name cells <<< "$step"
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${name}_${cells}.csv $PY bench.py --workload $name --cells $cells --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_${name}.log 2>&1; tail -3 gpurun_out/launches_${name}_${cells}.csv ;;
    abw:*)
      # abw:WORKLOAD:CELLS:V1,V2,... -- A/B of kernel variants on a slice of any workload
      IFS=: read -r _ name cells vs <<< "$step"
      for v in $(echo "$vs" | tr ',' ' '); do
        timeout 300 $PY bench.py --workload $name --cells $cells --steps 3 --warmup 3 --variant $v --no-cpu-baseline --no-e2e > gpurun_out/abw_${name}_${cells}_v$v.json 2> gpurun_out/abw_${name}_v$v.err
        $PY -c "import json; d=json.load(open('gpurun_out/abw_${name}_${cells}_v$v.json')); print('$name[$cells] variant $v: %.2f k solves/s, %.3f ms/step, k=%.2f, model frac %.3f, %s' % (d['value']/1e3, d['ms_per_step'], d['config']['mean_pcg_iterations'], d['roofline']['frac'], d['clocks']['reasons']))" || tail -3 gpurun_out/abw_${name}_v$v.err
      done ;;
    *) echo "unknown step $step" ;;
  esac
done

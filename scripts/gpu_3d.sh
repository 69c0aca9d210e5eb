#!/bin/bash
# 3D regression + bench + launch list (one GPU).  Usage: gpurun -- bash scripts/gpu_3d.sh
python -m pytest tests/test_gpu_parity_3d.py -q -x 2>&1 | tail -3
python bench.py --workload 3d-16x16 --steps 3 --no-cpu-baseline > gpurun_out/b3d_16.json 2> gpurun_out/b3d_16.err; tail -c 400 gpurun_out/b3d_16.err
python bench.py --workload 3d-8x32 --steps 3 --no-cpu-baseline > gpurun_out/b3d_32.json 2> gpurun_out/b3d_32.err; tail -c 400 gpurun_out/b3d_32.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_3d.csv python bench.py --workload 3d-16x16 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
echo done

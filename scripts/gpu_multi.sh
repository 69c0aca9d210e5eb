#!/bin/bash
# usage: gpu_multi.sh N   -- bench at 1..N GPUs of one box, both arms
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
cut -c1-330 gpurun_out/scale_n1.json; tail -2 gpurun_out/scale_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    grep '"metric"' gpurun_out/scale_n$n.json | cut -c1-330; tail -2 gpurun_out/scale_n$n.err
    grep -o '"e2e": {[^}]*}' gpurun_out/scale_n$n.json | cut -c1-200
  fi
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/scale_ref.json 2> gpurun_out/scale_ref.err
cut -c1-200 gpurun_out/scale_ref.json; tail -2 gpurun_out/scale_ref.err
# other BASELINE configurations and the 3D workload on all N GPUs
if [ "${EXTRA:-1}" = "1" ]; then
  for w in cfg3 cfg4 cfg5 3d-16x16; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 3 --warmup 3 --workload $w > gpurun_out/scale_${w}_n$N.json 2> gpurun_out/scale_${w}_n$N.err
    grep '"metric"' gpurun_out/scale_${w}_n$N.json | cut -c1-200; tail -1 gpurun_out/scale_${w}_n$N.err
  done
  python -m pytest tests/test_gpu_host_driver.py -q -k "two_gpus" 2>&1 | tail -2
fi

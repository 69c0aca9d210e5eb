#!/bin/bash
# last check of the tree: full GPU suite, smoke, default bench + reference arm, cluster scheduling policy A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
for pol in 0 1; do
  MSB_CLUSTER_POLICY=$pol timeout 200 python bench.py --workload cfg1x64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/pol${pol}_cfg1x64.json 2> gpurun_out/pol${pol}.err
  MSB_CLUSTER_POLICY=$pol timeout 200 python bench.py --workload cfg1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/pol${pol}_cfg1.json 2>> gpurun_out/pol${pol}.err
  python - <<PY
import json
for w in ("cfg1x64","cfg1"):
    try:
        d=json.load(open("gpurun_out/pol${pol}_%s.json"%w))
        print("policy $pol", w, "solves/s %.0f"%d["value"], "solve_ms %.3f"%d["roofline"]["kernel_ms_per_launch"])
    except Exception as e:
        print("policy $pol", w, "FAILED", e)
PY
done
timeout 600 python bench.py > gpurun_out/bench_target.json 2> gpurun_out/bench_target.err; cut -c1-160 gpurun_out/bench_target.json; tail -2 gpurun_out/bench_target.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-200 gpurun_out/bench_reference.json

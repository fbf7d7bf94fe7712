#!/bin/bash
# Standard GPU visit: parity tests (one process per file), tcgen05 probe, bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in post conv forward; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu -x --timeout 600 > gpurun_out/test_$f.log 2>&1
  echo "test_gpu_$f exit $?"; tail -5 gpurun_out/test_$f.log
done
timeout 900 python tools/tc_probe.py > gpurun_out/tc_probe.log 2>&1; cat gpurun_out/tc_probe.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err

#!/bin/bash
# Round-2 GPU visit: every -m gpu test file in its own process (a fault cannot poison the others), the queued tests, the bench with its
# extras.  Logs -> gpurun_out/.  Usage: bash tools/gpu_visit.sh [tests] [queued] [bench] [ncu] [dropin]
mkdir -p gpurun_out
what="${*:-tests queued dropin bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run_file() {
  local f=$1; shift
  timeout 1500 env "$@" python -m pytest tests/$f.py -q -m gpu --timeout 1200 > gpurun_out/$f.log 2>&1
  echo "$f exit $?: $(tail -1 gpurun_out/$f.log)"
  grep -E "^(FAILED|ERROR)" gpurun_out/$f.log | head -20
}
if [[ $what == *tests* ]]; then
  for f in test_gpu_conv test_gpu_post test_gpu_forward test_prep test_coco_format test_visualizer test_gpu_eager_bar; do run_file $f X=1; done
fi
if [[ $what == *queued* ]]; then run_file test_gpu_queued ORIENMASK_B200_QUEUED=1; fi
if [[ $what == *dropin* ]]; then run_file test_gpu_reference_dropin X=1; fi
if [[ $what == *bench* ]]; then
  timeout 1200 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 6000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
fi
if [[ $what == *ncu* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"; tail -2 gpurun_out/ncu_launches.log
fi

#!/bin/bash
# Standard GPU visit: parity tests (one process per file), bench, ncu launch list + one full capture.
# Logs -> gpurun_out/.  Usage: bash tools/gpu_round.sh [tests] [queued] [probe] [bench] [events] [ncu] [traffic] [full] [abtest] [fullx]
mkdir -p gpurun_out
what="${*:-tests bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
if [[ $what == *tests* ]]; then
  for f in post conv forward prep coco_format; do
    timeout 900 python -m pytest $( [[ $f == prep || $f == coco_format ]] && echo tests/test_$f.py || echo tests/test_gpu_$f.py ) -q -m gpu -x --timeout 600 > gpurun_out/test_$f.log 2>&1
    echo "test_gpu_$f exit $?"; tail -4 gpurun_out/test_$f.log
  done
fi
if [[ $what == *queued* ]]; then
  # tests written without a GPU (tests/test_gpu_queued.py): opt-in, one process so that a fault cannot poison the regular files
  ORIENMASK_B200_QUEUED=1 timeout 900 python -m pytest tests/test_gpu_queued.py -q -m gpu --timeout 600 > gpurun_out/test_queued.log 2>&1
  echo "test_gpu_queued exit $?"; tail -15 gpurun_out/test_queued.log
fi
if [[ $what == *probe* ]]; then
  timeout 900 python tools/tc_probe.py > gpurun_out/tc_probe.log 2>&1; cat gpurun_out/tc_probe.log
fi
if [[ $what == *bench* ]]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_ref.log
fi
if [[ $what == *events* ]]; then
  timeout 600 python tools/profile_step.py --steps 1 --events 10 > gpurun_out/events.log 2>&1; echo "events exit $?"; tail -2 gpurun_out/events.log
fi
if [[ $what == *ncu* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"; tail -2 gpurun_out/ncu_launches.log
fi
if [[ $what == *traffic* ]]; then
  # DRAM bytes of every launch of one step (one pass, 3 metrics) -> tools/traffic_report.py -> profiles/
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv \
      --log-file gpurun_out/traffic.csv python tools/profile_step.py --steps 1 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
fi
if [[ $what == *full* ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${CONV_KERNEL:-conv_tc2_kernel} -s ${CONV_SKIP:-84} -c ${CONV_COUNT:-1} -f -o gpurun_out/prof_conv \
      python tools/profile_step.py --steps 2 > gpurun_out/ncu_full_conv.log 2>&1; echo "ncu full conv exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mask_kernel -c 1 -f -o gpurun_out/prof_mask \
      python tools/profile_step.py --steps 2 > gpurun_out/ncu_full_mask.log 2>&1; echo "ncu full mask exit $?"
fi
if [[ $what == *abtest* ]]; then
  # A/B of an environment switch: AB_VAR=name AB_VALUES="0 1"
  for v in ${AB_VALUES:-0 1}; do
    env ${AB_VAR:-ORIENMASK_B200_PDL}=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$v.log 2>&1
    echo "${AB_VAR:-ORIENMASK_B200_PDL}=$v: $(python -c "import json,sys; d=json.loads(open('gpurun_out/bench_ab_$v.log').read().strip().splitlines()[-1]); print(round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms', d['roofline']['kernel'][-40:])" 2>&1 | tail -1)"
  done
fi
if [[ $what == *fullx* ]]; then
  # extra full captures: a flat-tile (im2col) layer, the RLE kernel, the pre-process kernel
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc2_kernel -s 44 -c 1 -f -o gpurun_out/prof_conv_flat \
      python tools/profile_step.py --steps 2 > gpurun_out/ncu_full_flat.log 2>&1; echo "ncu full flat exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mask_rle_kernel -c 1 -f -o gpurun_out/prof_rle \
      python tools/profile_step.py --steps 2 > gpurun_out/ncu_full_rle.log 2>&1; echo "ncu full rle exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:prep_kernel -c 1 -f -o gpurun_out/prof_prep \
      python tools/profile_step.py --steps 2 > gpurun_out/ncu_full_prep.log 2>&1; echo "ncu full prep exit $?"
fi

#!/bin/bash
# Round-2 evidence visit: every GPU test, the bench (+ reference arm), ncu launch lists / traffic / full captures, sanitizer.  -> gpurun_out/
mkdir -p gpurun_out
what="${*:-tests bench timeline ncu full sanitizer}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
if [[ $what == *tests* ]]; then
  for f in test_gpu_conv test_gpu_post test_gpu_forward test_prep test_coco_format test_visualizer test_gpu_eager_bar test_gpu_reference_dropin; do
    timeout 1500 python -m pytest tests/$f.py -q -m gpu --timeout 1200 > gpurun_out/$f.log 2>&1; echo "$f: $(tail -1 gpurun_out/$f.log)"; grep -E "^(FAILED|ERROR)" gpurun_out/$f.log | head
  done
  timeout 900 python -m pytest tests/test_gpu_shapes.py -q -m gpu > gpurun_out/test_gpu_shapes.log 2>&1; echo "shapes: $(tail -1 gpurun_out/test_gpu_shapes.log)"
fi
if [[ $what == *bench* ]]; then
  timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>> gpurun_out/bench.err; tail -c 600 gpurun_out/bench_ref.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_10steps.log 2>> gpurun_out/bench.err
  timeout 600 python bench.py --precision parity --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_parity.log 2>> gpurun_out/bench.err
fi
if [[ $what == *ncu* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 700 --csv --log-file gpurun_out/launches_bench_step.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --ncu-range > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench exit $?"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"; cp gpurun_out/layers.json gpurun_out/layers_fp16.json
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv \
      --log-file gpurun_out/traffic.csv python tools/profile_step.py --steps 1 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches_parity.csv \
      python tools/profile_step.py --steps 1 --precision parity --coco 0 > gpurun_out/ncu_launches_parity.log 2>&1; echo "ncu parity launches exit $?"
  timeout 600 python tools/profile_step.py --steps 1 --events 10 --coco 0 > /dev/null 2>&1; cp gpurun_out/layer_events.json gpurun_out/layer_events_fp16.json
fi
if [[ $what == *timeline* ]]; then
  timeout 300 python tools/timeline.py --md gpurun_out/timeline.md > gpurun_out/timeline.log 2>&1; tail -1 gpurun_out/timeline.log
  timeout 300 python tools/timeline.py --batch 1 --md gpurun_out/timeline_bs1.md > gpurun_out/timeline_bs1.log 2>&1; tail -1 gpurun_out/timeline_bs1.log
  timeout 300 python tools/timeline.py --precision parity --md gpurun_out/timeline_parity.md > gpurun_out/timeline_parity.log 2>&1; tail -1 gpurun_out/timeline_parity.log
  timeout 300 python tools/timeline.py --variants unfused:ORIENMASK_B200_FUSED_STEM=0,ORIENMASK_B200_FUSED_BLOCK=0 fused: --md gpurun_out/timeline_fusion_ab.md > /dev/null 2>&1
  timeout 120 python tools/phase_log.py > gpurun_out/phase_log_stem_fused.txt 2>&1
fi
if [[ $what == *full* ]]; then
  # conv_tc2 launch index = layer index - 2 in the fp16 schedule (fused stem + conv2.0, fused block), layer index - 1 in the parity one
  for spec in "conv_tc2_kernel 81 fp16 prof_conv136" "conv_tc2_kernel 84 parity prof_conv136_parity" "conv_tc2_kernel 43 fp16 prof_conv17_flat" \
              "conv_tc2_kernel 7 fp16 prof_conv68_res" "stem_fused_kernel 0 fp16 prof_stem_fused" "dark_block_kernel 0 fp16 prof_block" "stem_tc_kernel 0 parity prof_stem"; do
    set -- $spec
    timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$1 -s $2 -c 1 -f -o gpurun_out/$4 \
        python tools/profile_step.py --steps 2 --precision $3 --coco 0 > gpurun_out/ncu_$4.log 2>&1; echo "ncu full $4 exit $?"
  done
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"mask_kernel|conf_compact|nms_kernel|select_tail" -c 4 -f -o gpurun_out/prof_post \
      python tools/profile_step.py --steps 2 --coco 0 > gpurun_out/ncu_prof_post.log 2>&1; echo "ncu full post exit $?"
fi
if [[ $what == *sanitizer* ]]; then
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_conv.py -q -m gpu -x -k "split_precision_engine or head_channel or parity_split_layouts" > gpurun_out/sanitizer.log 2>&1
  echo "sanitizer exit $?: $(grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer.log | tail -2 | tr '\n' ' ')"
  timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_forward.py -q -m gpu -x -k "c_engine or fp16_small or parity_small" > gpurun_out/sanitizer_fwd.log 2>&1
  echo "sanitizer (forward: C engine, fused stem, fused block) exit $?: $(grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_fwd.log | tail -2 | tr '\n' ' ')"
fi

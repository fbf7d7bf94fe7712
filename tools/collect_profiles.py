"""Copy the outputs of tools/gpu_evidence.sh from gpurun_out/ (scratch) into profiles/r02_* (tracked), deriving the summaries:
launch list of the timed bench steps per kernel, per-layer tables, DRAM traffic, ncu full-capture summary, sanitizer summary."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def cp(a, b):
    if os.path.exists(os.path.join(G, a)):
        shutil.copy(os.path.join(G, a), os.path.join(P, b))


def bench_line(src, dst):
    path = os.path.join(G, src)
    if not os.path.exists(path):
        return
    lines = [l for l in open(path).read().splitlines() if l.startswith('{')]
    if lines:
        json.dump(json.loads(lines[-1]), open(os.path.join(P, dst), 'w'), indent=1)


for a, b in (('parity_e2e_parity.json', 'r02_parity_e2e_parity.json'), ('parity_e2e_fp32.json', 'r02_parity_e2e_fp32.json'),
             ('parity_e2e_plain_init.json', 'r02_parity_e2e_plain_init.json'), ('parity_stage_plain_init.json', 'r02_parity_stage_plain_init.json'),
             ('dropin_infer_py_parity.json', 'r02_dropin_infer_py_parity.json'), ('dropin_infer_py_fp16.json', 'r02_dropin_infer_py_fp16.json'),
             ('dropin_test_py_parity.json', 'r02_dropin_test_py_parity.json'), ('reference_gpu_bar.json', 'r02_reference_gpu_bar.json'),
             ('drift.json', 'r02_forward_drift.json'), ('fp16_agreement.json', 'r02_fp16_detection_agreement.json'),
             ('fp16_error_on_trained_like_heads.json', 'r02_fp16_error_on_trained_like_heads.json'), ('eager_bar.json', 'r02_eager_pytorch_bar.json'),
             ('launches.csv', 'r02_launches_step.csv'), ('launches_parity.csv', 'r02_launches_step_parity.csv'),
             ('timeline.md', 'r02_timeline.md'), ('timeline_bs1.md', 'r02_timeline_bs1.md'), ('timeline_parity.md', 'r02_timeline_parity.md'),
             ('timeline_fusion_ab.md', 'r02_timeline_fusion_ab.md'), ('phase_log_stem_fused.txt', 'r02_phase_log_stem_fused.txt'),
             ('tl_ab3.md', 'r02_timeline_l2hint_ab.md'), ('tl_ab1.md', 'r02_timeline_resdirect_ab.md')):
    cp(a, b)
for a, b in (('bench.log', 'r02_bench_1gpu.json'), ('bench_10steps.log', 'r02_bench_1gpu_10steps.json'), ('bench_parity.log', 'r02_bench_1gpu_parity.json'),
             ('bench_ref.log', 'r02_bench_reference_arm.json'), ('bench_2gpu.log', 'r02_bench_2gpu.json'), ('bench_4gpu.log', 'r02_bench_4gpu.json'),
             ('bench_8gpu.log', 'r02_bench_8gpu.json')):
    bench_line(a, b)
reps = [os.path.join(G, n) for n in ('prof_conv136.ncu-rep', 'prof_conv136_parity.ncu-rep', 'prof_conv17_flat.ncu-rep', 'prof_conv68_res.ncu-rep',
                                      'prof_stem_fused.ncu-rep', 'prof_block.ncu-rep', 'prof_stem.ncu-rep', 'prof_post.ncu-rep')
        if os.path.exists(os.path.join(G, n))]
if reps:
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_summary.py')] + reps, capture_output=True, text=True).stdout
    open(os.path.join(P, 'r02_ncu_summary.txt'), 'w').write(out)
san = os.path.join(G, 'sanitizer.log')
if os.path.exists(san):
    keep = [l for l in open(san).read().splitlines() if 'SUMMARY' in l or 'passed' in l or 'COMPUTE-SANITIZER' in l]
    extra = os.path.join(G, 'sanitizer_fwd.log')
    keep2 = [l for l in open(extra).read().splitlines() if 'SUMMARY' in l or 'passed' in l] if os.path.exists(extra) else []
    open(os.path.join(P, 'r02_sanitizer.txt'), 'w').write(
        'compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_conv.py -q -m gpu -x -k "split_precision_engine or head_channel or parity_split_layouts"\n'
        + '\n'.join(keep) + '\n\ncompute-sanitizer --tool memcheck python -m pytest tests/test_gpu_forward.py -q -m gpu -x -k "c_engine or fp16_small or parity_small"'
        '   (C engine, fused stem + conv2.0, fused block, TMA stem, all precisions)\n' + '\n'.join(keep2) + '\n')
path = os.path.join(G, 'launches_bench_step.csv')
if os.path.exists(path):
    rows, tot = collections.OrderedDict(), 0.0
    for r in csv.DictReader([l for l in open(path) if not l.startswith('==')]):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        us = float(r['Metric Value'].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r.get('Metric Unit', 'ns'), 1e-3)
        k = r['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        a = rows.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
    with open(os.path.join(P, 'r02_launches_bench_step.csv'), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --ncu-range\n'
                '# the launches of the 2 TIMED steps (cudaProfilerStart/Stop around the device-resident loop); times are cold-cache and serialised: use the shares\n'
                'kernel,launches,us,share\n')
        for k, (n, us) in rows.items():
            f.write('%s,%d,%.1f,%.4f\n' % (k, n, us, us / tot))
    print(open(os.path.join(P, 'r02_launches_bench_step.csv')).read())
for src, layers, dst in (('layer_events_fp16.json', 'layers_fp16.json', 'r02_layers_events.md'), ('launches.csv', 'layers_fp16.json', 'r02_layers_ncu.md')):
    if os.path.exists(os.path.join(G, src)):
        subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'layer_report.py'), os.path.join(G, src), '--layers', os.path.join(G, layers),
                        '--md', os.path.join(P, dst)], capture_output=True)
if os.path.exists(os.path.join(G, 'traffic.csv')):
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'traffic_report.py'), os.path.join(G, 'traffic.csv'), '--json', os.path.join(P, 'r02_traffic.json')],
                   capture_output=True)

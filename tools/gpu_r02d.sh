#!/bin/bash
# r02 third single-GPU pass: push-tile with global reds, pull with 8 entries per lane; full test-suite on every layout.
mkdir -p gpurun_out
timeout 900 python tools/kernel_ab.py c2 --reps 12 --variants "soa;tiles:SC=global;tiles:SC=global,RB=14;pull;pull:RB=14;pull:EPL=4;pull:EPL=4,RB=14" > gpurun_out/ab3_c2.log 2>&1
echo "ab c2 exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/ab3_c2.log'):
    try:
        d = json.loads(l); print(d['workload'][:12], d['variant'], 'cold', round(d.get('kernel_us_cold_median', -1), 1), 'min', round(d.get('kernel_us_cold_min', -1), 1), 'warm', round(d.get('kernel_us_warm_median', -1), 1), 'frac', round(d.get('frac_cold', 0), 3), d.get('grad_max_diff_over_max'), d.get('loss_rel_diff_vs_first'), d.get('error'))
    except Exception as e: print('bad', l[:300])
PY
timeout 900 python tools/kernel_ab.py c3 c5 --variants "soa;tiles:SC=global,RB=14;pull;pull:RB=14;pull:EPL=4" > gpurun_out/ab3_large.log 2>&1
echo "ab large exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/ab3_large.log'):
    try:
        d = json.loads(l); print(d['workload'][:12], d['variant'], 'cold', round(d.get('kernel_us_cold_median', -1), 1), 'frac', round(d.get('frac_cold', 0), 3), d.get('grad_max_diff_over_max'), d.get('loss_rel_diff_vs_first'), d.get('error'))
    except Exception as e: print('bad', l[:300])
PY
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --timeout=800 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
MDE_B200_LAYOUT=pull timeout 600 ncu --set full --clock-control none --import-source on -k regex:distortion_pull -s 3 -c 1 -o gpurun_out/r02_pull8_c2 -f python tools/prof_target.py 8 > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_full.log
timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1200; tail -3 gpurun_out/bench.err

#!/bin/bash
# r02: hybrid pull (repulsive edges as push entries) A/B + parity on every layout
mkdir -p gpurun_out
timeout 900 python tools/kernel_ab.py c2 m134 --reps 12 --variants "soa;pull;pull:REP=mirror;pull:EPL=8;pull:RB=14" > gpurun_out/ab6_small.log 2>&1
echo "ab exit $?"; python - <<'PY'
import json
for l in open('gpurun_out/ab6_small.log'):
    try:
        d = json.loads(l); print(d['workload'][:12], d['variant'], 'cold', round(d.get('kernel_us_cold_median', -1), 1), 'min', round(d.get('kernel_us_cold_min', -1), 1), 'warm', round(d.get('kernel_us_warm_median', -1), 1), 'frac', round(d.get('frac_cold', 0), 3), d.get('grad_max_diff_over_max'), d.get('loss_rel_diff_vs_first'), d.get('error'))
    except Exception as e: print('bad', l[:300])
PY
timeout 900 python tools/kernel_ab.py c5 --variants "soa;pull;pull:REP=mirror" > gpurun_out/ab6_c5.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/ab6_c5.log'):
    try:
        d = json.loads(l); print(d['workload'][:12], d['variant'], 'cold', round(d.get('kernel_us_cold_median', -1), 1), 'frac', round(d.get('frac_cold', 0), 3), d.get('grad_max_diff_over_max'), d.get('loss_rel_diff_vs_first'), d.get('error'))
    except Exception as e: print('bad', l[:300])
PY
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_state.py tests/test_gpu_configs.py -m gpu -q --maxfail=20 --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
MDE_B200_LAYOUT=pull timeout 600 ncu --set full --clock-control none --import-source on -k regex:distortion_pull -s 3 -c 1 -o gpurun_out/r02_pullhybrid_c2 -f python tools/prof_target.py 8 > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -1 gpurun_out/ncu_full.log

#!/bin/bash
# r02 second GPU pass: pull kernel parity + A/B against the SoA / push-tile kernels, ncu capture.
mkdir -p gpurun_out
timeout 600 python tools/kernel_ab.py c2 m134 --reps 12 --variants "soa;pull;pull:RB=14;pull:RB=12" > gpurun_out/ab2_small.log 2>&1
echo "ab_small exit $?"; grep -E '"C2"|error' gpurun_out/ab2_small.log | cut -c1-330
grep -E 'm[134] ' gpurun_out/ab2_small.log | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d['workload'], d['variant'], round(d.get('kernel_us_cold_median', -1), 1), d.get('grad_max_diff_over_max'), d.get('loss_rel_diff_vs_first'), d.get('error'))
    except Exception as e: print('bad', l[:200])
"
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/kernel_ab.py c3 c5 --variants "soa;tiles;pull;pull:RB=14" > gpurun_out/ab2_large.log 2>&1
echo "ab_large exit $?"; python -c "
import sys, json
for l in open('gpurun_out/ab2_large.log'):
    try:
        d = json.loads(l); print(d['workload'][:12], d['variant'], round(d.get('kernel_us_cold_median', -1), 1), round(d.get('frac_cold', 0), 3), d.get('grad_max_diff_over_max'), d.get('loss_rel_diff_vs_first'), d.get('error'))
    except Exception as e: print('bad', l[:200])
"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -x -k "golden or zero or widths" --timeout=500 > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit $?"; tail -3 gpurun_out/sanitizer.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:distortion_pull -s 3 -c 1 -o gpurun_out/r02_pull_c2 -f python tools/prof_target.py 8 > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_full.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-3000; tail -3 gpurun_out/bench.err

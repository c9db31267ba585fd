#!/bin/bash
# r02 fourth single-GPU pass: full GPU test-suite (all layouts), solver launch lists, bench.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python tools/kernel_ab.py c3 --variants "soa;pull;pull:EPL=4" > gpurun_out/ab4_c3.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/ab4_c3.log'):
    try:
        d = json.loads(l); print(d['workload'][:12], d['variant'], 'cold', round(d.get('kernel_us_cold_median', -1), 1), 'frac', round(d.get('frac_cold', 0), 3), d.get('grad_max_diff_over_max'), d.get('loss_rel_diff_vs_first'), d.get('error'))
    except Exception as e: print('bad', l[:300])
PY
PYMDE_B200_SOLVER_MODE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/r02_launches_warm_hoststep.csv python tools/prof_target.py 12 > gpurun_out/prof_launch.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 900 --csv --log-file gpurun_out/r02_launches_warm_stepgraph.csv python tools/prof_target.py 12 >> gpurun_out/prof_launch.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_warm_hoststep.csv 2>/dev/null | head -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgs_dots -s 8 -c 1 -o gpurun_out/r02_lbfgs_dots -f python tools/prof_target.py 14 > gpurun_out/ncu_dots.log 2>&1
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-1500; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-1200

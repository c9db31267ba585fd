#!/bin/bash
# r02 final single-GPU pass: full GPU suite, smoke, C5-shard ncu captures (SoA and pull), final bench + reference arm.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:distortion_quad -s 2 -c 1 -o gpurun_out/r02_c5shard_soa -f python tools/prof_c5.py > gpurun_out/ncu_c5_soa.log 2>&1
echo "ncu soa exit $?"; tail -1 gpurun_out/ncu_c5_soa.log
MDE_B200_LAYOUT=pull timeout 900 ncu --set full --clock-control none --import-source on -k regex:distortion_pull -s 2 -c 1 -o gpurun_out/r02_c5shard_pull -f python tools/prof_c5.py > gpurun_out/ncu_c5_pull.log 2>&1
echo "ncu pull exit $?"; tail -1 gpurun_out/ncu_c5_pull.log
MDE_B200_DETERMINISTIC=1 timeout 600 python tools/kernel_ab.py c2 --reps 8 --variants "soa" > gpurun_out/ab5_det.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/ab5_det.log'):
    try:
        d = json.loads(l); print('deterministic', d['workload'][:12], d['variant'], 'cold', round(d.get('kernel_us_cold_median', -1), 1), d.get('error'))
    except Exception as e: print('bad', l[:300])
PY
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?"; tail -1 gpurun_out/bench.log | cut -c1-900; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 100 --warmup 5 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400

#!/bin/bash
# r02: ELL pull kernel v2 (descriptor + pre-issued tile, K lane-slots per lane, balanced CTA ranges, rsqrt distance):
# full GPU suite, A/B, bench with the ELL layout forced, ncu capture, smoke and the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -20 | cut -c1-250
timeout 300 python tools/kernel_ab.py c2 --reps 12 --variants 'soa;ell;ell:PACK=0' > gpurun_out/ab_ell2_c2.jsonl 2> gpurun_out/ab_ell2_c2.err
echo "ab c2 exit $?"; cut -c1-330 gpurun_out/ab_ell2_c2.jsonl; tail -3 gpurun_out/ab_ell2_c2.err
timeout 400 python tools/kernel_ab.py c3 --reps 6 --variants 'pull;ell' > gpurun_out/ab_ell2_c3.jsonl 2> gpurun_out/ab_ell2_c3.err
echo "ab c3 exit $?"; cut -c1-330 gpurun_out/ab_ell2_c3.jsonl; tail -3 gpurun_out/ab_ell2_c3.err
MDE_B200_LAYOUT=ell timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ell.json 2> gpurun_out/bench_ell.err
echo "bench ell exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ell.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("ms_per_step", "iters_per_sec")}, d["e2e"]["iters_per_sec"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
except Exception as ex:
    print("bench ell parse", ex)
PY
MDE_B200_LAYOUT=ell ncu --set full --clock-control none --import-source on -k regex:distortion_ell -s 4 -c 1 -o gpurun_out/r02_ell2_c2 -f python tools/prof_target.py 8 > gpurun_out/prof_full_ell.log 2>&1
echo "full capture exit $?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err; tail -c 1800 gpurun_out/bench.log | head -c 900; tail -3 gpurun_out/bench.err

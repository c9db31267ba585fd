#!/bin/bash
# r02: ELL pull kernel (mde_ell.cu) -- parity tests on the new layout, A/B against soa / pull at C2, C3-shaped, m = 1/3/4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=300 -k "ell" > gpurun_out/pytest_ell.log 2>&1
rc=$?
echo "pytest ell exit $rc"; tail -5 gpurun_out/pytest_ell.log | cut -c1-300
timeout 200 python tools/ell_tiny.py > gpurun_out/ell_tiny.log 2>&1
echo "tiny exit $?"; tail -12 gpurun_out/ell_tiny.log | cut -c1-300
if [ $rc -ne 0 ]; then
  grep -E "Error|error|assert|Mismatch|mismatch" gpurun_out/pytest_ell.log | head -20 | cut -c1-300
  timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/ell_tiny.py > gpurun_out/ell_sanitizer.log 2>&1
  echo "sanitizer exit $?"; grep -E "Invalid|ERROR SUMMARY|at 0x|by thread" gpurun_out/ell_sanitizer.log | head -30 | cut -c1-300
  exit 1
fi
timeout 300 python tools/kernel_ab.py c2 --reps 12 --variants 'soa;pull:EPL=4;ell;ell:RB=12' > gpurun_out/ab_ell_c2.jsonl 2> gpurun_out/ab_ell_c2.err
echo "ab c2 exit $?"; cut -c1-420 gpurun_out/ab_ell_c2.jsonl; tail -3 gpurun_out/ab_ell_c2.err
timeout 400 python tools/kernel_ab.py c3 m134 --reps 6 --variants 'soa;pull;ell' > gpurun_out/ab_ell_large.jsonl 2> gpurun_out/ab_ell_large.err
echo "ab large exit $?"; cut -c1-420 gpurun_out/ab_ell_large.jsonl; tail -3 gpurun_out/ab_ell_large.err
MDE_B200_LAYOUT=ell timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ell.json 2> gpurun_out/bench_ell.err
echo "bench ell exit $?"; cut -c1-1500 gpurun_out/bench_ell.json; tail -3 gpurun_out/bench_ell.err
MDE_B200_LAYOUT=ell ncu --set full --clock-control none --import-source on -k regex:distortion_ell -s 4 -c 1 -o gpurun_out/r02_ell_c2 -f python tools/prof_target.py 8 > gpurun_out/prof_full_ell.log 2>&1
echo "full capture exit $?"
MDE_B200_LAYOUT=ell ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches_warm_ell.csv python tools/prof_target.py 12 > gpurun_out/prof_launch_ell.log 2>&1
echo "warm list exit $?"

#!/bin/bash
# r02 final pass: device-built ELL records (bit-compared with the host builder), full GPU suite, smoke, bench (C2 line +
# roofline_c3), ncu capture of the ELL kernel at the C3 shape
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ell_build.py -m gpu -q -x --timeout=200 > gpurun_out/pytest_ell_build.log 2>&1
echo "pytest ell build exit $?"; tail -6 gpurun_out/pytest_ell_build.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -20 | cut -c1-250
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "iters_per_sec")}, "e2e", d["e2e"]["iters_per_sec"])
    print("roofline", d["roofline"]["kernel_ms"], d["roofline"]["frac"])
    print("roofline_c3", d.get("roofline_c3"))
except Exception as ex:
    print("bench parse", ex)
PY
tail -3 gpurun_out/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:distortion_ell -s 3 -c 1 -o gpurun_out/r02_ell_c3 -f python tools/kernel_ab.py c3 --reps 4 --variants 'ell' > gpurun_out/prof_full_ell_c3.log 2>&1
echo "full capture c3 exit $?"; tail -2 gpurun_out/prof_full_ell_c3.log | cut -c1-400

#!/bin/bash
# r02 last pass: branch-free Huber, guard-free coefficient for functions with finite f'(0), unmasked full rows in the ELL
# kernel -- kernel / config / builder / solver tests, then the bench line (C2 + roofline_c3)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs.py tests/test_gpu_ell_build.py tests/test_gpu_solver.py tests/test_gpu_state.py -m gpu -q --maxfail=10 --timeout=300 > gpurun_out/pytest_last.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_last.log; tail -4 gpurun_out/pytest_last.log | cut -c1-300
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_last.log | head -20 | cut -c1-250
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_last.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "iters_per_sec")}, "e2e", d["e2e"]["iters_per_sec"], sorted(d["e2e"]["seconds_per_call"])[:3])
    print("roofline", d["roofline"]["kernel_ms"], d["roofline"]["frac"])
    print("roofline_c3", d.get("roofline_c3"))
except Exception as ex:
    print("bench parse", ex)
PY
tail -3 gpurun_out/bench_last.err

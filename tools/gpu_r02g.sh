#!/bin/bash
# r02: f4 tests (spectral / BFS), full GPU suite; then (2 GPUs) the sharded bench with the reworked peer loads
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_setup.py tests/test_gpu_multi.py tests/test_gpu_configs.py -m gpu -q --timeout=800 > gpurun_out/pytest_setup.log 2>&1
echo "pytest setup exit $?"; tail -30 gpurun_out/pytest_setup.log | cut -c1-400
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_n2.log').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'timed_windows_ms')})
    print('single', d.get('single_gpu')); print('c2', d.get('c2_sharded')); print('parity', {k: v for k, v in d.get('parity', {}).items() if k != 'what'})
except Exception as e:
    print('no line', e); print(open('gpurun_out/bench_n2.err').read()[-1500:])
PY

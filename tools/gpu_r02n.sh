#!/bin/bash
# r02: 2 GPUs with the late-epilogue step chain: parity tests, then the sharded bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout=500 > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi exit $?"; tail -6 gpurun_out/pytest_multi.log | cut -c1-300
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

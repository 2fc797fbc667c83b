#!/bin/bash
# strips leg only, N ranks (development runs)
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --strips-only > gpurun_out/r2f_strips_n$N.json 2> gpurun_out/r2f_strips_n$N.err || tail -20 gpurun_out/r2f_strips_n$N.err
python - <<PY
import json
r = json.loads(open('gpurun_out/r2f_strips_n$N.json').read().strip().splitlines()[-1])['strips']
print({k: v for k, v in r.items() if k not in ('per_rank_pass_ms', 'exchange', 'config', 'one_gpu_pass_ms', 'balancing')})
print('one GPU', r.get('one_gpu_pass_ms'))
print('balancing', r.get('balancing'))
for i, p in enumerate(r.get('per_rank_pass_ms') or []):
    print('rank', i, 'sum %.3f' % sum(p.values()), p)
PY

#!/bin/bash
# Round 2, GPU run E (N GPUs, N = $1): strips bit-identity over real NVLink peers (N = 2 only) and the bench at N ranks (replicas + strips leg).
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi topo -m > gpurun_out/r2e_topo_n$N.txt 2>&1
if [ "$N" = "2" ]; then
  echo "== pytest strips (2 GPUs)"; ( time timeout 600 python -m pytest tests/test_strips_gpu.py -m gpu -q ) > gpurun_out/r2e_pytest_n2.txt 2>&1; tail -4 gpurun_out/r2e_pytest_n2.txt
fi
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 40 --warmup 4 > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err || tail -20 gpurun_out/r2e_bench_n$N.err
python - <<PY
import json
try:
    r = json.loads(open('gpurun_out/r2e_bench_n$N.json').read().strip().splitlines()[-1])
    print('N=$N value %.1f ms %.4f e2e %.1f (%.4f ms)' % (r['value'], r['ms_per_step'], r['e2e']['value'], r['e2e']['ms_per_step']))
    print("strips", json.dumps(r.get("strips"))[:3000])
    print('issue', r['config'].get('issue'), r['config'].get('host_affinity'))
except Exception as e:
    print('unreadable', e)
PY

#!/bin/bash
# Round 2, run K1: where the Bloom cluster tail should start (per-level launches vs the single cluster launch), finer up-sampling chunks.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== levels kernel test"; timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "bloom" 2>&1 | tail -3
run() { # name, DFX_TUNE, extra bench flags
  DFX_TUNE="$2" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-psnr --no-strips $3 > gpurun_out/r2k1_$1.json 2> gpurun_out/r2k1_$1.err || tail -3 gpurun_out/r2k1_$1.err
}
run default ""
run tail_off "bloom_tail=0"
run tail_8k "bloom_tail_texels=8192"
run tail_32k "bloom_tail_texels=32768"
run rows2 "bloom_up_rows=2"
run rows3 "bloom_up_rows=3"
run default_b ""
run tail_off_b "bloom_tail=0"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2k1_*.json')):
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    p = {x['pass']: x['ms'] for x in r['passes']}
    b = sum(v for k, v in p.items() if k.startswith('bloom'))
    print('%-12s step %.4f e2e %.4f launches %d bloom %.4f  ' % (f.split('r2k1_')[1][:-5], r['ms_per_step'], r['e2e']['ms_per_step'], r['gpu_launches'], b) +
          ' '.join('%s=%.4f' % (k[6:], v) for k, v in p.items() if k.startswith('bloom')))
PY

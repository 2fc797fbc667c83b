#!/bin/bash
# Round 2, run I: A/B of the Bloom up-sampling launch shapes (rows per warp, pipeline depth, operator as template parameter), parity of each.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { # name, DFX_TUNE
  DFX_TUNE="$2" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-psnr --no-strips > gpurun_out/r2i_$1.json 2> gpurun_out/r2i_$1.err || tail -3 gpurun_out/r2i_$1.err
}
run default ""
run tm_rt "bloom_up_tm_template=0"
run deep "bloom_up_deep=1"
run wave "bloom_up_rows=0"
run rows4 "bloom_up_rows=4"
run rows16 "bloom_up_rows=16"
run wave_deep "bloom_up_rows=0,bloom_up_deep=1"
run rows16_deep "bloom_up_rows=16,bloom_up_deep=1"
run rows32 "bloom_up_rows=32"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2i_*.json')):
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    p = {x['pass']: x['ms'] for x in r['passes']}
    print('%-12s step %.4f e2e %.4f  up %.4f composite %.4f  prefilter %.4f down %.4f tail %.4f' % (f.split('r2i_')[1][:-5], r['ms_per_step'], r['e2e']['ms_per_step'],
          p.get('bloom_upsample', -1), p.get('bloom_composite_tonemap', -1), p.get('bloom_prefilter', -1), p.get('bloom_downsample', -1), p.get('bloom_tail', -1)))
PY
echo "== parity (default)"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_fullsize_gpu.py -q -m gpu -k "bloom or fused or chain" 2>&1 | tail -3
echo "== parity (deep, run-time operator)"; DFX_TUNE="bloom_up_deep=1,bloom_up_tm_template=0,bloom_up_rows=0" timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "bloom or fused" 2>&1 | tail -3

#!/bin/bash
# Round 2, run K1b: do the cluster tails of the depth pyramids cost the overlapped frame what the Bloom tail did? (Bloom tail now off by default.)
mkdir -p gpurun_out
export PYTHONPATH=$PWD
run() { # name, DFX_TUNE, extra bench flags
  DFX_TUNE="$2" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-psnr --no-strips $3 > gpurun_out/r2k1b_$1.json 2> gpurun_out/r2k1b_$1.err || tail -3 gpurun_out/r2k1b_$1.err
}
run default ""
run pyrtail_off "pyramid_tail=0"
run pyr_perlevel "pyramid_impl=0"
run default_b ""
run pyrtail_off_b "pyramid_tail=0"
run pyr_perlevel_b "pyramid_impl=0"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2k1b_*.json')):
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    p = {x['pass']: x['ms'] for x in r['passes']}
    print('%-16s step %.4f e2e %.4f launches %d serial sum %.4f  ' % (f.split('r2k1b_')[1][:-5], r['ms_per_step'], r['e2e']['ms_per_step'], r['gpu_launches'], sum(p.values())) +
          ' '.join('%s=%.4f' % (k, p[k]) for k in ('ssr_hiz', 'ssao_prefilter_depth', 'ssao_convolute')))
PY

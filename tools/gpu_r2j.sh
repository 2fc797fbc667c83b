#!/bin/bash
# Round 2, run J: the single-launch Bloom levels kernel (cooperative, grid-wide barriers) - bit-identity test, A/B in the bench with and
# without async compute, chain parity with it switched on, ncu --set full of the Bloom kernels of both forms.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== levels kernel test"; timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "bloom" 2>&1 | tail -15
run() { # name, DFX_TUNE, extra bench flags
  DFX_TUNE="$2" timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-psnr --no-strips $3 > gpurun_out/r2j_$1.json 2> gpurun_out/r2j_$1.err || tail -3 gpurun_out/r2j_$1.err
}
run perlevel ""
run levels "bloom_levels=1"
run perlevel_1stream "" "--no-overlap"
run levels_1stream "bloom_levels=1" "--no-overlap"
run levels_nograph "bloom_levels=1" "--no-graph"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2j_*.json')):
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    p = {x['pass']: x['ms'] for x in r['passes']}
    print('%-18s step %.4f e2e %.4f launches %d  ' % (f.split('r2j_')[1][:-5], r['ms_per_step'], r['e2e']['ms_per_step'], r['gpu_launches']) +
          ' '.join('%s=%.4f' % (k[6:], v) for k, v in p.items() if k.startswith('bloom')))
PY
echo "== chain parity with the levels kernel"; DFX_TUNE="bloom_levels=1" timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_fullsize_gpu.py tests/test_cpp_shim.py -q -m gpu -k "bloom or fused or chain or async or stream or shim" 2>&1 | tail -4
echo "== ncu full (Bloom kernels)"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"bloom_" -c 9 -o gpurun_out/r2j_bloom_perlevel python tools/ncu_target.py > gpurun_out/r2j_ncu_a.log 2>&1; tail -1 gpurun_out/r2j_ncu_a.log
DFX_TUNE="bloom_levels=1" timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"bloom_" -c 3 -o gpurun_out/r2j_bloom_levels python tools/ncu_target.py > gpurun_out/r2j_ncu_b.log 2>&1; tail -1 gpurun_out/r2j_ncu_b.log
for n in perlevel levels; do ncu -i gpurun_out/r2j_bloom_$n.ncu-rep --page raw --csv > gpurun_out/r2j_ncu_bloom_$n.csv 2> /dev/null; done
ls -la gpurun_out/*.ncu-rep; rm -f gpurun_out/*.ncu-rep; du -sh gpurun_out

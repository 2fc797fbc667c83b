#!/bin/bash
# Round 2, GPU run B: new Bloom (warp-shuffle streaming + cluster tail), TMA pyramid kernels, native chain executor + CUDA graphs.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/r2b_smoke.txt
echo "== pytest"; ( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2b_pytest.txt 2>&1; tail -8 gpurun_out/r2b_pytest.txt
run() { name=$1; shift; echo "== bench $name"; timeout 400 "$@" > gpurun_out/r2b_bench_$name.json 2> gpurun_out/r2b_bench_$name.err || tail -5 gpurun_out/r2b_bench_$name.err; }
run default python bench.py
Q="--no-cpu-baseline --no-psnr --steps 40"
DFX_TUNE=bloom_impl=0,bloom_tail=0 run bloom_r1 python bench.py $Q
DFX_TUNE=ssao_spatial_impl=0 run spatial_gather python bench.py $Q
DFX_TUNE=pyramid_impl=0 run pyr_levels python bench.py $Q
DFX_TUNE=pyramid_impl=1 run pyr_ldg python bench.py $Q
run nograph python bench.py $Q --no-graph
run nooverlap python bench.py $Q --no-overlap
run 1080p python bench.py $Q --width 1920 --height 1080
run 1080p_nograph python bench.py $Q --width 1920 --height 1080 --no-graph
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_*.json')):
    try:
        r = json.loads(open(f).read())
    except Exception as e:
        print(f, 'unreadable', e); continue
    p = {x['pass']: x['ms'] for x in r['passes']}
    live = {x['pass']: x.get('live', {}).get('ms', -1) for x in r['passes']}
    print('   live: spatial %.4f resample %.4f temporal %.4f' % (live.get('ssao_spatial', -1), live.get('ssao_resample', -1), live.get('ssao_temporal', -1)))
    keys = ['ssr_hiz', 'ssao_prefilter_depth', 'ssao_convolute', 'bloom_prefilter', 'bloom_downsample', 'bloom_tail', 'bloom_upsample', 'bloom_composite_tonemap']
    print(f.split('r2b_bench_')[1][:-5].ljust(14), 'step %.4f e2e %.4f launches %d' % (r['ms_per_step'], r['e2e']['ms_per_step'], r['gpu_launches']), ' '.join('%s=%.4f' % (k.replace('bloom_', 'b_').replace('ssao_', 'a_'), p.get(k, -1)) for k in keys), r['config'].get('issue', {}).get('frames_replayed'), r.get('psnr', {}) and r['psnr'].get('ldr'))
PY
# launch list (eager issue: one kernel node per launch), cold-cache serialised times: compare shares
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 120 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-psnr --no-graph --no-overlap > /dev/null 2>&1
tail -3 gpurun_out/r2b_launches.csv

#!/bin/bash
# Round 2, GPU run C: native-format inputs, one-wave Bloom streaming, strips executor (virtual ranks), ncu --set full of the kernels the
# north star names, compute-sanitizer over the smoke frame.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest"; ( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2c_pytest.txt 2>&1; tail -12 gpurun_out/r2c_pytest.txt
run() { name=$1; shift; echo "== bench $name"; timeout 500 "$@" > gpurun_out/r2c_bench_$name.json 2> gpurun_out/r2c_bench_$name.err || tail -5 gpurun_out/r2c_bench_$name.err; }
run default python bench.py --steps 40
Q="--no-cpu-baseline --no-psnr --steps 40"
run nograph python bench.py $Q --no-graph
run 1080p python bench.py $Q --width 1920 --height 1080
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2c_bench_*.json')):
    try:
        r = json.loads(open(f).read())
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f.split('r2c_bench_')[1][:-5].ljust(10), 'step %.4f e2e %.4f launches %d' % (r['ms_per_step'], r['e2e']['ms_per_step'], r['gpu_launches']), r.get('psnr') and r['psnr'].get('ldr'), r['config'].get('issue'))
    for x in r['passes']:
        print('   %-26s %.4f frac %.3f  live %s' % (x['pass'], x['ms'], x['frac'], x.get('live')))
PY
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"bloom_|ssao_|ssr_intersect|pyramid_tile|taa_kernel|ssr_spatial|ssr_temporal|ssr_bilateral|postfx_prepare" -c 60 -o gpurun_out/r2c_full python tools/ncu_target.py > gpurun_out/r2c_ncu.log 2>&1; tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/r2c_full.ncu-rep
echo "== compute-sanitizer"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2c_memcheck.log 2>&1; tail -4 gpurun_out/r2c_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r2c_racecheck.log 2>&1; tail -4 gpurun_out/r2c_racecheck.log

#!/bin/bash
# Round 2, final 1-GPU run (K2): the whole GPU suite, smoke, the default bench line + the reference arm, one-stream and graph-less comparisons,
# the launch list of the same command.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest"; ( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2k2_pytest.txt 2>&1; tail -6 gpurun_out/r2k2_pytest.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2k2_smoke.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/r2k2_bench_n1.json 2> gpurun_out/r2k2_bench_n1.err || tail -5 gpurun_out/r2k2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k2_bench_reference.json 2> gpurun_out/r2k2_bench_reference.err
Q="--steps 20 --warmup 4 --no-cpu-baseline --no-psnr --no-strips"
timeout 300 python bench.py $Q --no-overlap > gpurun_out/r2k2_bench_1stream.json 2> /dev/null
timeout 300 python bench.py $Q --no-graph > gpurun_out/r2k2_bench_nograph.json 2> /dev/null
DFX_TUNE="pyramid_impl=2" timeout 300 python bench.py $Q > gpurun_out/r2k2_bench_pyr_tile.json 2> /dev/null
python - <<'PY'
import json
r = json.loads(open('gpurun_out/r2k2_bench_n1.json').read().strip().splitlines()[-1])
print('value %.1f (%.4f ms) e2e %.1f (%.4f ms, %.1f GB/s) launches %d psnr %s' % (r['value'], r['ms_per_step'], r['e2e']['value'], r['e2e']['ms_per_step'], r['e2e']['h2d_GBps'], r['gpu_launches'], r.get('psnr')))
print('roofline', r['roofline']); print('cpu', r['cpu_baseline'])
for x in r['passes']:
    print('   %-26s %.4f frac %.3f share %.3f live %s' % (x['pass'], x['ms'], x['frac'], x['share'], x.get('live')))
print(open('gpurun_out/r2k2_bench_reference.json').read()[:400])
for n in ('1stream', 'nograph', 'pyr_tile'):
    try:
        q = json.loads(open('gpurun_out/r2k2_bench_%s.json' % n).read().strip().splitlines()[-1]); print(n, q['ms_per_step'], q['e2e']['ms_per_step'], q['gpu_launches'])
    except Exception as e:
        print(n, 'unreadable', e)
PY
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 140 --csv --log-file gpurun_out/r2k2_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-psnr --no-strips --no-graph --no-overlap > /dev/null 2>&1
tail -2 gpurun_out/r2k2_launches.csv | cut -c1-200

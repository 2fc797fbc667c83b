#!/bin/bash
# Round 2, final 1-GPU run: the whole GPU suite, smoke, the default bench line, the launch list of the same command.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest"; ( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2h_pytest.txt 2>&1; tail -6 gpurun_out/r2h_pytest.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2h_smoke.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err || tail -5 gpurun_out/r2h_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/r2h_bench_n1.json').read().strip().splitlines()[-1])
print('value %.1f (%.4f ms) e2e %.1f (%.4f ms, %.1f GB/s) launches %d psnr %s' % (r['value'], r['ms_per_step'], r['e2e']['value'], r['e2e']['ms_per_step'], r['e2e']['h2d_GBps'], r['gpu_launches'], r.get('psnr')))
print('roofline', r['roofline'])
for x in r['passes']:
    print('   %-26s %.4f frac %.3f share %.3f live %s' % (x['pass'], x['ms'], x['frac'], x['share'], x.get('live')))
print(open('gpurun_out/r2h_bench_reference.json').read()[:600])
PY
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-psnr --no-graph --no-overlap > /dev/null 2>&1
tail -2 gpurun_out/r2h_launches.csv | cut -c1-200

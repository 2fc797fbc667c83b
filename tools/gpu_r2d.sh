#!/bin/bash
# Round 2, GPU run D (1 GPU): G-buffer format A/B for the resident arm, ncu --set full of the blur / Bloom kernels (converted to CSV on the box).
mkdir -p gpurun_out
export PYTHONPATH=$PWD
echo "== pytest (fast subset)"; ( time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_strips_gpu.py tests/test_cpp_shim.py -m gpu -q ) > gpurun_out/r2d_pytest.txt 2>&1; tail -4 gpurun_out/r2d_pytest.txt
run() { name=$1; shift; echo "== bench $name"; timeout 500 "$@" > gpurun_out/r2d_bench_$name.json 2> gpurun_out/r2d_bench_$name.err || tail -5 gpurun_out/r2d_bench_$name.err; }
Q="--no-cpu-baseline --no-psnr --steps 60"
run native python bench.py $Q
run fp32 python bench.py $Q --gbuffer fp32
run native_b python bench.py $Q
run fp32_b python bench.py $Q --gbuffer fp32
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2d_bench_*.json')):
    try:
        r = json.loads(open(f).read())
    except Exception as e:
        print(f, 'unreadable', e); continue
    p = {x['pass']: x['ms'] for x in r['passes']}
    print(f.split('r2d_bench_')[1][:-5].ljust(10), 'step %.4f e2e %.4f' % (r['ms_per_step'], r['e2e']['ms_per_step']), ' '.join('%s=%.4f' % (k[:14], v) for k, v in p.items()))
PY
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"bloom_up2x_stream|bloom_down2x_stream|ssao_spatial|ssao_resample|ssao_temporal|taa_kernel|ssr_intersect|ssr_bilateral|ssao_ao" -c 36 -o gpurun_out/r2d_full python tools/ncu_target.py > gpurun_out/r2d_ncu.log 2>&1; tail -2 gpurun_out/r2d_ncu.log
ncu -i gpurun_out/r2d_full.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_raw.csv 2> /dev/null
ls -la gpurun_out/r2d_full.ncu-rep gpurun_out/r2d_ncu_raw.csv
[ $(stat -c %s gpurun_out/r2d_full.ncu-rep) -gt 45000000 ] && rm gpurun_out/r2d_full.ncu-rep
rm -f gpurun_out/r2c_full.ncu-rep
du -sh gpurun_out

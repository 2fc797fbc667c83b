#!/bin/bash
# Occupancy sweep of the latency-bound kernels: rebuild with -DDFX_OCC_*=N and print the per-pass times (run on the GPU box).
for n in 3 4 5 6; do
  DFX_NVCC_EXTRA="-DDFX_OCC_INTERSECT=$n -DDFX_OCC_SSR_SPATIAL=$n -DDFX_OCC_SSR_TEMPORAL=$n -DDFX_OCC_AO=$n -DDFX_OCC_TAA=$n" python -m diligentfx_b200.build --force > /dev/null 2>&1
  python bench.py --steps 40 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read())
p={x['pass']:x['ms'] for x in r['passes']}
print('occ $n step %.3f ms | intersect %.4f spatial %.4f temporal %.4f ao %.4f taa %.4f' % (r['ms_per_step'], p['ssr_intersect'], p['ssr_spatial'], p['ssr_temporal'], p['ssao_ambient_occlusion'], p.get('taa', p.get('compose_taa', 0.0))))"
done
python -m diligentfx_b200.build --force > /dev/null 2>&1

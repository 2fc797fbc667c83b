#!/bin/bash
# Times bench.py with pre-built variants of the library (diligentfx_b200/lib/variants/*.so, built here with
# DFX_NVCC_EXTRA="-DDFX_OCC_…=N" python -m diligentfx_b200.build --force) — run on the GPU box. DFX_LIB selects the build.
for v in default "$@"; do
  lib=""; [ "$v" != default ] && lib="diligentfx_b200/lib/variants/$v.so"
  DFX_LIB="$lib" python bench.py --steps 60 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read()); p={x['pass']:x['ms'] for x in r['passes']}
print('%-8s step %.4f ms | intersect %.4f spatial %.4f temporal %.4f ao %.4f compose_taa %.4f' % ('$v', r['ms_per_step'], p['ssr_intersect'], p['ssr_spatial'], p['ssr_temporal'], p['ssao_ambient_occlusion'], p['compose_taa']))"
done

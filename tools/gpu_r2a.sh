#!/bin/bash
# Round 2, GPU run A: full-size parity tests, baseline bench with the PSNR leg, the opt-in variants built in the container.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2a_smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.txt 2>&1
tail -5 gpurun_out/r2a_pytest.txt
timeout 600 python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_err.txt
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r2a_bench_n1.json').read())
print('bench', r['value'], r['ms_per_step'], 'e2e', r['e2e']['value'], 'psnr', r.get('psnr'))
PY
cat > /tmp/tma_frames.py <<'PY'
import hashlib
import numpy as np
from diligentfx_b200 import synth
from diligentfx_b200.chain import ChainConfig, PostProcessChain
for (w, h) in ((1920, 1080), (256, 144), (130, 70)):
    seq = synth.generate_sequence(w, h, 3)
    chain = PostProcessChain(w, h, ChainConfig())
    for fr in seq:
        ldr = chain.run_frame(fr).cpu().numpy()
    print(w, h, hashlib.sha256(ldr.tobytes()).hexdigest()[:16], bool(np.isfinite(ldr).all()))
PY
echo "== default";   timeout 300 python /tmp/tma_frames.py | tee gpurun_out/r2a_tma_default.txt
echo "== bloom_tma"; DFX_LIB=diligentfx_b200/lib/variants/bloom_tma.so timeout 300 python /tmp/tma_frames.py | tee gpurun_out/r2a_tma_variant.txt
cmp gpurun_out/r2a_tma_default.txt gpurun_out/r2a_tma_variant.txt && echo "TMA variant: frames bit-identical to the default kernels" | tee -a gpurun_out/r2a_tma_variant.txt
timeout 600 bash tools/variant_sweep.sh bloom_tma intersect_v2 intersect_v2_occ5 2>&1 | tee gpurun_out/r2a_variants.txt

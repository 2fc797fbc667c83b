#!/bin/bash
# Round-2 first GPU run for the opt-in TMA variant of the Bloom 2:1 down-sampling: build it beside the
# default library, check that the chain's frames are bit-identical to the default kernels, run the Bloom parity tests on it,
# then time both. Run on the GPU box (nvcc rebuilds there: `python -m diligentfx_b200.build --force` takes about a minute).
set -e
mkdir -p diligentfx_b200/lib/variants gpurun_out
DFX_NVCC_EXTRA="-DDFX_BLOOM_TMA=1" python -m diligentfx_b200.build --force > /dev/null
cp diligentfx_b200/lib/libdfx_b200.so diligentfx_b200/lib/variants/bloom_tma.so
python -m diligentfx_b200.build --force > /dev/null
cat > /tmp/tma_frames.py <<'PY'
import hashlib, sys
import numpy as np
from diligentfx_b200 import synth
from diligentfx_b200.chain import ChainConfig, PostProcessChain
for (w, h) in ((3840, 2160), (1920, 1080), (256, 144), (130, 70)):
    seq = synth.generate_sequence(w, h, 3)
    chain = PostProcessChain(w, h, ChainConfig())
    for fr in seq:
        ldr = chain.run_frame(fr).cpu().numpy()
    print(w, h, hashlib.sha256(ldr.tobytes()).hexdigest()[:16], bool(np.isfinite(ldr).all()))
PY
echo "== default";   timeout 600 python /tmp/tma_frames.py | tee gpurun_out/tma_default.txt
echo "== bloom_tma"; DFX_LIB=diligentfx_b200/lib/variants/bloom_tma.so timeout 600 python /tmp/tma_frames.py | tee gpurun_out/tma_variant.txt
cmp gpurun_out/tma_default.txt gpurun_out/tma_variant.txt && echo "TMA variant: frames bit-identical to the default kernels"
DFX_LIB=diligentfx_b200/lib/variants/bloom_tma.so timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "bloom or fused or full_chain"
bash tools/variant_sweep.sh bloom_tma

#!/bin/bash
# Builds the opt-in kernel variants beside the default library (diligentfx_b200/lib/variants/<name>.so), for
# tools/variant_sweep.sh <name>... on the GPU box. Timed in round 2 (profiles/r2a): intersect_v2 within 1.5 % of the default.
#   bloom_tma          -DDFX_BLOOM_TMA=1                         TMA staging of the Bloom 2:1 down-sampling tile
#   intersect_v2       -DDFX_INTERSECT_V2=1                      Hi-Z march loop, 40 instead of 42 instructions per step (spills 40 B at 40 registers)
#   intersect_v2_occ5  -DDFX_INTERSECT_V2=1 -DDFX_OCC_INTERSECT=5  the same with 48 registers (no spills, 5 CTAs / SM)
set -e
mkdir -p diligentfx_b200/lib/variants
build() { DFX_NVCC_EXTRA="$2" python -m diligentfx_b200.build --force > /dev/null && cp diligentfx_b200/lib/libdfx_b200.so "diligentfx_b200/lib/variants/$1.so" && echo "built $1 ($2)"; }
build bloom_tma "-DDFX_BLOOM_TMA=1"
build intersect_v2 "-DDFX_INTERSECT_V2=1"
build intersect_v2_occ5 "-DDFX_INTERSECT_V2=1 -DDFX_OCC_INTERSECT=5"
python -m diligentfx_b200.build --force > /dev/null
echo "default library restored; next: bash tools/variant_sweep.sh bloom_tma intersect_v2 intersect_v2_occ5"

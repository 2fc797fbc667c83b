#!/bin/bash
# strips: bit-identity tests (virtual ranks + 2 real GPUs) and the strips-only bench leg
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( time timeout 600 python -m pytest tests/test_strips_gpu.py -m gpu -q ) > gpurun_out/r2g_pytest_n$N.txt 2>&1; tail -4 gpurun_out/r2g_pytest_n$N.txt
bash tools/gpu_r2f.sh $N

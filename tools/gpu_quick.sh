#!/bin/bash
# quick GPU pass: parity tests + variant sweeps (no ncu).  gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag>'
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 $OUT/pytest_gpu.log
for kb in 64 32 24; do echo "D3M_GATHER_SMEM_KB=$kb"; D3M_GATHER_SMEM_KB=$kb timeout 300 python tools/bp_variants.py "" ; done > $OUT/bp_variants.txt 2>&1 ; cat $OUT/bp_variants.txt

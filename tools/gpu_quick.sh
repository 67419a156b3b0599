#!/bin/bash
# quick GPU pass: parity tests + variant sweeps (no ncu).  gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag>'
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/bp_variants.py "" "24:3:2,40:5:2,80:5:4" "24:2:3,40:2:5,80:4:5" > $OUT/bp_variants.txt 2>&1 ; cat $OUT/bp_variants.txt
timeout 300 python tools/tsdf_variants.py 0 1 2 3 > $OUT/tsdf_variants.txt 2>&1 ; cat $OUT/tsdf_variants.txt

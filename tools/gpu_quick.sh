#!/bin/bash
# quick GPU pass: parity tests + environment-variant sweep of the headline step (no ncu).
#   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag>'
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -3 $OUT/pytest_gpu.log
D3M_FWD_TVMIN=4 D3M_FWD_KU=4 D3M_FWD_WARPS_PER_SM=32 timeout 300 python -m pytest tests/test_gpu_back_project.py tests/test_gpu_shard.py -m gpu -x -q > $OUT/pytest_gpu_variant.log 2>&1 ; echo "pytest(variant) rc=$?" ; tail -3 $OUT/pytest_gpu_variant.log
D3M_PDL=0 timeout 300 python -m pytest tests/test_gpu_back_project.py -m gpu -x -q > $OUT/pytest_gpu_nopdl.log 2>&1 ; echo "pytest(no pdl) rc=$?" ; tail -2 $OUT/pytest_gpu_nopdl.log
timeout 600 python tools/step_variants.py "" "D3M_PDL=0" "D3M_FWD_STAGE_KR=0" "D3M_FWD_TVMIN=4" "D3M_FWD_KU=4" \
  "D3M_FWD_WARPS_PER_SM=32" "D3M_FWD_TVMIN=4,D3M_FWD_WARPS_PER_SM=32" "D3M_FWD_TVMIN=4,D3M_FWD_WARPS_PER_SM=32,D3M_FWD_KU=4" \
  "D3M_FWD_TVMIN=4,D3M_FWD_KU=4" > $OUT/step_variants.txt 2> $OUT/step_variants.err ; echo "variants rc=$?"
cat $OUT/step_variants.txt ; tail -5 $OUT/step_variants.err

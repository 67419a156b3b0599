#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu --set full of the hot kernels.
# Run as:  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -5 $OUT/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -2 $OUT/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err ; echo "bench rc=$?" ; tail -3 $OUT/bench.err ; cat $OUT/bench.json
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err ; echo "ref rc=$?" ; cat $OUT/bench_ref.json
echo "== ncu launch list (bp step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_bp.csv python bench.py --profile-step bp --steps 2 --warmup 3 --no-graph > $OUT/ncu_bp.log 2>&1 ; echo "rc=$?"
echo "== ncu launch list (dense + tsdf)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_dense.csv python bench.py --profile-step dense --steps 1 --warmup 3 --no-graph > $OUT/ncu_dense.log 2>&1 ; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_tsdf.csv python bench.py --profile-step tsdf --steps 1 --warmup 3 --no-graph > $OUT/ncu_tsdf.log 2>&1 ; echo "rc=$?"
echo "== ncu full: dense level-2 fwd + bwd kernels"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $OUT/prof_dense -f python bench.py --profile-step dense --steps 1 --warmup 3 --no-graph > $OUT/ncu_full_dense.log 2>&1 ; echo "rc=$?"
echo "== ncu full: headline fragment step (roofline.traffic)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $OUT/prof_bp -f python bench.py --profile-step bp --steps 1 --warmup 3 --no-graph > $OUT/ncu_full_bp.log 2>&1 ; echo "rc=$?"
echo "== ncu full: tsdf"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tsdf_integrate -c 2 -o $OUT/prof_tsdf -f python bench.py --profile-step tsdf --steps 1 --warmup 3 --no-graph > $OUT/ncu_full_tsdf.log 2>&1 ; echo "rc=$?"
ls -la $OUT
echo "== tsdf host profile" ; timeout 300 python tools/tsdf_host_profile.py > $OUT/tsdf_host_profile.txt 2>&1 ; cat $OUT/tsdf_host_profile.txt

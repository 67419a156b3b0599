#!/bin/bash
# One GPU-box pass of round 2.  gpurun --timeout 1500 -- 'bash tools/gpu_pass.sh <tag> [tests|bench|launches|ncu|sanitize ...]'
TAG=${1:-r02}; shift
WHAT=${@:-tests bench launches}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
for w in $WHAT; do
case $w in
tests) timeout 1100 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error" $OUT/pytest_gpu.log | tail -3; grep -E "^(FAILED|ERROR|E  )" $OUT/pytest_gpu.log | head -30;;
smoke) timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log;;
bench) timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -4 $OUT/bench.err;;
benchq) timeout 600 python bench.py --steps 20 --warmup 3 --no-reference-gpu --no-reference-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -4 $OUT/bench.err;;
ref) timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?";;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_bp.csv python bench.py --profile-step bp --steps 2 --warmup 3 --no-graph > $OUT/ncu_bp.log 2>&1; echo "launches bp rc=$?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_dense.csv python bench.py --profile-step dense --steps 1 --warmup 3 --no-graph > $OUT/ncu_dense.log 2>&1; echo "launches dense rc=$?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_tsdf.csv python bench.py --profile-step tsdf --steps 1 --warmup 3 --no-graph > $OUT/ncu_tsdf.log 2>&1; echo "launches tsdf rc=$?";;
ncu)
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $OUT/prof_bp -f python bench.py --profile-step bp --steps 1 --warmup 3 --no-graph > $OUT/ncu_full_bp.log 2>&1; echo "ncu bp rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $OUT/prof_dense -f python bench.py --profile-step dense --steps 1 --warmup 3 --no-graph > $OUT/ncu_full_dense.log 2>&1; echo "ncu dense rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tsdf_integrate -c 2 -o $OUT/prof_tsdf -f python bench.py --profile-step tsdf --steps 1 --warmup 3 --no-graph > $OUT/ncu_full_tsdf.log 2>&1; echo "ncu tsdf rc=$?";;
sanitize)
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_back_project.py tests/test_gpu_tsdf.py -m gpu -x -q -k "not large_scene and not dense_levels" > $OUT/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/sanitizer_memcheck.txt
  timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_back_project.py tests/test_gpu_tsdf.py -m gpu -x -q -k "not large_scene and not dense_levels and not crowded" > $OUT/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/sanitizer_racecheck.txt;;
esac
done
ls $OUT

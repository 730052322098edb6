#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list and one full capture of the top kernel.
# Usage (from the build container):  gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > $OUT/gpu.csv 2>&1
free -g > $OUT/host.txt; nproc >> $OUT/host.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee $OUT/bench_reference.json
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -3 | tee $OUT/bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 --sustained-steps 0 > $OUT/ncu_launches_bench.log 2>&1
tail -3 $OUT/launches.csv
echo "== ncu full (stft kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-stft2048} -s 3 -c 1 -f -o $OUT/prof_stft \
    python bench.py --steps 3 --warmup 3 --clips 128 --no-cpu --e2e-steps 0 --sustained-steps 0 > $OUT/ncu_full_bench.log 2>&1
ls -la $OUT
fi

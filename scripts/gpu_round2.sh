#!/bin/bash
# One gpurun call of round 2: GPU parity tests (all, no -x), smoke, the headline bench with every config.
# Usage:  gpurun --timeout 1500 -- 'bash scripts/gpu_round2.sh TAG'
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > $OUT/gpu.csv 2>&1
free -g > $OUT/host.txt; nproc >> $OUT/host.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu ${PYTEST_ARGS:-} 2>&1 | tail -40 | tee $OUT/pytest_gpu.log
fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
if [ "${SKIP_BENCH:-0}" != "1" ]; then
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee $OUT/bench_reference.json
echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:-} 2> $OUT/bench.err | tail -3 | tee $OUT/bench.json; tail -20 $OUT/bench.err
fi

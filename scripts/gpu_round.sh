#!/bin/bash
# One gpurun call: GPU parity tests, every-config device-resident bench, the headline bench.
# Usage:  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh TAG'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv > $OUT/gpu.csv 2>&1
free -g > $OUT/host.txt; nproc >> $OUT/host.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu ${PYTEST_ARGS:-} 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== configs"; timeout 900 python scripts/bench_configs.py --out $OUT/configs.jsonl ${CONFIG_ARGS:-} 2>&1 | tail -30 | tee $OUT/configs.log
if [ "${SKIP_BENCH:-0}" != "1" ]; then
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee $OUT/bench_reference.json
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -3 | tee $OUT/bench.json
fi

#!/bin/bash
# ncu full capture of one launch of each kernel matching $NCU_KERNEL (regex) while running bench_configs.
# Usage: gpurun --timeout 900 -- 'NCU_KERNEL=mdct2048 CONFIG_ARGS="--only mdct --scale 0.125 --steps 2" bash scripts/gpu_ncu.sh TAG'
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 800 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-2048_warp} -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-1} -f -o $OUT/prof_${NCU_NAME:-k} \
    python scripts/bench_configs.py ${CONFIG_ARGS:-} > $OUT/ncu_${NCU_NAME:-k}.log 2>&1
tail -3 $OUT/ncu_${NCU_NAME:-k}.log
ls -la $OUT

#!/bin/bash
# One iteration on the GPU: a pytest -k selection, a bench_configs --only selection, optionally one ncu capture.
#   gpurun -- 'K="cqt or mel" ONLY=cqt,mfcc REGEX=cqt_eo_kernel NCU_ONLY=cqt bash scripts/gpu_iter.sh TAG'
TAG=${1:-iter}
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ -n "$K" ]; then timeout 900 python -m pytest tests -q -m gpu -k "$K" 2>&1 | tail -25 | tee $OUT/pytest.log; fi
if [ -n "$ONLY" ]; then timeout 600 python scripts/bench_configs.py --only $ONLY --out $OUT/configs.jsonl 2>&1 | cut -c1-330 | tail -12; fi
if [ -n "$REGEX" ]; then bash scripts/ncu_kernel.sh $REGEX ${NCU_ONLY:-cqt} ${TAG}_ncu 2>&1 | tail -32; fi

#!/bin/bash
# CQT iteration on one GPU: the cqt parity tests, the cfg-5 timing on both routes, one ncu capture of the even/odd kernel.
TAG=${1:-cqt}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -k "cqt" 2>&1 | tail -15 | tee $OUT/pytest.log
timeout 600 python scripts/bench_configs.py --only cqt,cqttc --out $OUT/configs.jsonl 2>&1 | tail -5
if [ "${NCU:-1}" = "1" ]; then
  REGEX=${REGEX:-cqt_eo_kernel} NCU_SKIP=${NCU_SKIP:-1} bash scripts/ncu_kernel.sh ${REGEX:-cqt_eo_kernel} cqt ${TAG}_ncu 2>&1 | tail -40
fi

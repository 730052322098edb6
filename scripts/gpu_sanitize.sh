#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over the small GPU parity tests.
# usage: gpurun --timeout 1800 -- 'bash scripts/gpu_sanitize.sh [TAG]'
TAG=${1:-sanitize}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SEL='not full_size and not two_rank and not million and not cfg3_batch and not cfg4 and not cfg2'
export ZAFB_PINNED_POOL_MB=0
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 \
      python -m pytest tests/test_gpu_stft.py tests/test_gpu_transforms.py -x -q -m gpu -k "$SEL and ${KSEL:-not tensor_core and not cqt_32768 and not golden}" \
      -p no:cacheprovider > $OUT/$tool.log 2>&1
  echo "exit $?" >> $OUT/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " $OUT/$tool.log | tail -5
done

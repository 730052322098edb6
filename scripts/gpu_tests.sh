#!/bin/bash
# GPU parity tests only (optionally under compute-sanitizer for the small cases).
TAG=${1:-t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu ${PYTEST_ARGS:-} 2>&1 | tail -60 | tee $OUT/pytest_gpu.log

#!/bin/bash
# C-order MDCT / IMDCT kernels: parity test, then cfg-4 timing of the kernel variants
OUT=gpurun_out/${1:-mdctbm}; mkdir -p $OUT
timeout 600 python -m pytest tests -q --tb=short -m gpu -k "bin_major or mdct" 2>&1 | grep -v "^E   " | tail -40 | tee $OUT/pytest.log
timeout 300 python scripts/probes/corder_probe.py ${SCALE:-1.0} mdct 2>&1 | grep mdct | tee $OUT/probe.txt


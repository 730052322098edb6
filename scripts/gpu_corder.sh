#!/bin/bash
# C-order kernels: parity tests, then cfg-2 / cfg-4 timing in the reference's memory order
OUT=gpurun_out/${1:-corder}; mkdir -p $OUT
timeout 600 python -m pytest tests -q --tb=short -m gpu -k "bin_major or mdct" 2>&1 | grep -v "^E   " | tail -40 | tee $OUT/pytest.log
timeout 300 python scripts/probes/corder_probe.py ${SCALE:-1.0} 2>&1 | grep transform | tee $OUT/probe.txt

#!/bin/bash
# Final record of round 2c: GPU tests, smoke, reference arm, bench, then the ncu launch list of the bench command.
bash scripts/gpu_round2.sh r02c_final
OUT=gpurun_out/r02c_final
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --sustained-steps 0 --e2e-steps 1 --no-cpu --config-steps 2 > $OUT/bench_under_ncu.log 2>&1
tail -c 300 $OUT/bench_under_ncu.log | head -3; wc -l $OUT/launches.csv

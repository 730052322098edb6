#!/bin/bash
# One `ncu --set full` capture of the first launch matching a kernel-name regex while scripts/bench_configs.py runs one
# transform; prints the summary and a few memory-system counters, keeps only the CSV.
# usage: gpurun -- 'bash scripts/ncu_kernel.sh transpose_tile stftbin [TAG]'   (CMD="python scripts/stft_probe.py 4096:1024:480000:128" overrides the workload)
REGEX=${1:?kernel regex}; ONLY=${2:?bench_configs --only value}; TAG=${3:-ncu_$ONLY}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s ${NCU_SKIP:-1} -c 1 -f -o $OUT/prof \
    ${CMD:-python scripts/bench_configs.py --only $ONLY --scale ${SCALE:-0.125} --steps 2} > $OUT/log.txt 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/source.csv 2>/dev/null
python scripts/ncu_summary.py < $OUT/raw.csv | tee $OUT/summary.txt
rm -f $OUT/prof.ncu-rep

#!/bin/bash
# One gpurun call: link probe, ncu launch list of the bench command, one `--set full` capture of
# every specialised kernel (raw CSV exported on the box; only the CSVs travel back).
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh TAG'
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== link probe"; timeout 300 python scripts/pcie_probe.py --out $OUT/pcie_probe.jsonl 2>&1 | head -20
echo "== ncu launch list (bench.py)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 --sustained-steps 0 > $OUT/ncu_launches_bench.log 2>&1
tail -3 $OUT/launches.csv
full() {  # name, kernel regex, skip, command...
    local name=$1 regex=$2 skip=$3; shift 3
    echo "== ncu full: $name"
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o $OUT/prof_$name \
        "$@" > $OUT/ncu_$name.log 2>&1
    if [ -f $OUT/prof_$name.ncu-rep ]; then
        ncu -i $OUT/prof_$name.ncu-rep --page raw --csv > $OUT/${name}_ncu_raw.csv 2>/dev/null
        python scripts/ncu_summary.py < $OUT/${name}_ncu_raw.csv | tee $OUT/${name}_ncu_summary.txt
        [ "$name" = "${KEEP_REP:-stft2048}" ] || rm -f $OUT/prof_$name.ncu-rep
    else
        tail -5 $OUT/ncu_$name.log
    fi
}
full stft2048 '^stft_warp_kernel' 3 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --sustained-steps 0
S="--scale 0.125 --steps 2"
full istft2048 '^istft_warp_kernel' 3 python scripts/bench_configs.py --only istft $S
full mdct2048 '^mdct_warp_kernel' 3 python scripts/bench_configs.py --only mdct $S
full imdct2048 '^imdct_warp_kernel' 3 python scripts/bench_configs.py --only imdct $S
full dct1024 dct1024_warp 3 python scripts/bench_configs.py --only dct $S
full stft_binmajor '^stft_warp_binmajor' 1 python scripts/bench_configs.py --only stftbin $S
full stft4096 '^stft_warp_kernel' 3 python scripts/stft_probe.py 4096:1024:480000:128
full istft4096 '^istft_warp_kernel' 3 python scripts/stft_probe.py 4096:1024:480000:128
full stft256 '^stft_warp_kernel' 3 python scripts/stft_probe.py 256:64:80000:512
full mdct512 '^mdct_warp_kernel' 3 python scripts/mdct_probe.py 512:1323000:256
full transpose transpose_tile 1 python scripts/mdct_probe.py 2048:1323000:64
for k in ${EXTRA_KERNELS:-}; do
    full $k $k 2 python scripts/bench_configs.py --only ${EXTRA_ONLY:-mel,mfcc,cqt,dct} $S
done
ls -la $OUT

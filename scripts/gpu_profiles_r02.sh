#!/bin/bash
# Round-2 evidence for profiles/: launch list of the bench command, one `ncu --set full` capture per kernel that changed.
OUT=gpurun_out/r02prof; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --sustained-steps 0 --e2e-steps 1 --no-cpu --config-steps 2 > $OUT/bench_under_ncu.log 2>&1
tail -c 600 $OUT/bench_under_ncu.log | head -5
NCU_SKIP=1 bash scripts/ncu_kernel.sh "stft_warp_kernel" stft r02prof_stft 2>&1 | tail -28
NCU_SKIP=1 bash scripts/ncu_kernel.sh cqt_eo_kernel cqt r02prof_cqt 2>&1 | tail -28
CMD="python scripts/probes/corder_probe.py 0.25" NCU_SKIP=1 bash scripts/ncu_kernel.sh istft_binmajor_kernel x r02prof_istft_bm 2>&1 | tail -28
SCALE=0.25 NCU_SKIP=1 bash scripts/ncu_kernel.sh gemm3xtf32_kernel dct r02prof_dct1 2>&1 | tail -28
SCALE=0.125 NCU_SKIP=0 bash scripts/ncu_kernel.sh mel_warp_kernel_f64 melf64 r02prof_melf64 2>&1 | tail -28

#!/bin/bash
# A/B runs of tuning switches (environment flags) on the device-resident config bench.
# usage: gpu_sweep.sh TAG "ENV1=a ENV2=b" "ENV1=c" ...     (ONLY=... selects the transforms)
TAG=${1:-sweep}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { echo "== $*"; env $* python scripts/bench_configs.py --only ${ONLY:-istft,mdct,imdct} --steps ${STEPS:-10} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('  %-16s %8.3f ms  %.3f of HBM peak  %.4g units/s' % (d['transform'], d['ms_per_step'], d['hbm_frac'], d['units_per_sec']))
"; }
for cfg in "$@"; do run $cfg; done | tee $OUT/sweep.log

fmt() { python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('  %-16s %8.3f ms  %.3f of HBM peak' % (d['transform'], d['ms_per_step'], d['hbm_frac']))
"; }
for b in 0 1; do echo "stft store-only bulk=$b"; ZAFB_STFT_PREFETCH=2 ZAFB_STFT_BULK=$b python scripts/bench_configs.py --only stft --steps 50 | fmt; done

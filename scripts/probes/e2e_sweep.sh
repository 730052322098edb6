#!/bin/bash
# Where the half-spectrum host path loses time: the same call with the fill disabled, fewer fill threads, other stage sizes.
run() { echo "== $*"; env "$@" python scripts/e2e_probe.py --chunks 16 2>&1 | grep "chunk 16 MB rep [12]"; }
run ZAFB_X=0
run ZAFB_HOST_MIRROR_NOFILL=1
run ZAFB_HOST_MIRROR_THREADS=8
run ZAFB_HOST_MIRROR_THREADS=12
run ZAFB_HOST_MIRROR=0
for mb in 8 32 64; do echo "== stage $mb MB"; python scripts/e2e_probe.py --chunks $mb 2>&1 | grep "rep [12]" | head -2; done

#!/bin/bash
# Static SASS evidence for profiles/: per-kernel instruction histograms of the shipped libzafb200.so, with the tcgen05 / TMEM /
# TMA mnemonics (UTCHMMA, LDTM, UTMALDG, UBLKCP, UTCBAR, SYNCS) counted explicitly.  Runs without a GPU.
#   bash scripts/sass_evidence.sh profiles/r02_sass
OUT=${1:-profiles/sass}
LIB=zaf-python_b200/libzafb200.so
TMP=$(mktemp)
cuobjdump -sass $LIB > $TMP
{
echo "# cuobjdump -sass $LIB ($(date -u +%Y-%m-%dT%H:%MZ)); static instruction counts per kernel"
echo "# kernels: $(grep -c 'Function :' $TMP)"
echo
echo "## tensor-core / TMEM / TMA mnemonics per kernel (only kernels that contain any)"
python3 - "$TMP" <<'PY'
import collections, re, sys
cur = None
hist = collections.defaultdict(collections.Counter)
for line in open(sys.argv[1]):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
keys = ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOM", "SYNCS")
for k, h in hist.items():
    found = {kk: sum(v for op, v in h.items() if op.startswith(kk)) for kk in keys}
    if any(found.values()):
        print(k[:110])
        print("    " + ", ".join(f"{kk} x{v}" for kk, v in found.items() if v))
PY
echo
for k in gemm3xtf32_pair_kernelILb1 gemm3xtf32_pair_kernelILb0 gemm3xtf32_kernelILi128 gemm3xtf32_kernelILi64 stft_warp_kernelILi2048ELb0ELi6ELb0 stft_warp_kernelILi2048ELb1ELi8 stft_warp_binmajor_kernelILi2048ELb0 \
         istft_warp_kernelILi2048ELi4ELi8ELb0 mdct_warp_kernelILi2048 imdct_warp_kernelILi2048 mel_warp_kernelILi1024ELi1ELb0 mel_warp_kernel_f64 cqt_eo_kernelILb0 cqt32768_kernelILb0 dct_warp_kernelILi1024ELi2ELb0; do
  echo "## $k"
  python3 scripts/sass_hist.py $k < $TMP 2>/dev/null | head -16
  echo
done
} > $OUT.txt
rm -f $TMP
wc -l $OUT.txt

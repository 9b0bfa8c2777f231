#!/bin/bash
# LumaEncoder::encode / LumaDecoder::decode through the reference's public API, same driver built against the facade
# (GPU) and against the unmodified reference sources (CPU), loopback codec: per-call times of 4K frames.
cd tests/cxx/build
for b in facade_roundtrip ref_roundtrip; do
  echo "== $b 3840x2160 PQ Lu'v' 11/8 profile 2, 5 frames"
  ./$b 3840 2160 1 0 11 8 2 12 1.0 5 2>&1 >/dev/null | grep "^time" | awk '{s[$2]+=$4; n[$2]++; if ($3>0) {t[$2]+=$4; m[$2]++}} END {for (k in s) printf "%s: mean %.2f ms over %d calls (%.2f ms without the first)\n", k, s[k]/n[k], n[k], t[k]/m[k]}'
done
echo "== facade_roundtrip with LUMA_FACADE_TIMING=1: transform vs run() per call (mean without the first call)"
LUMA_FACADE_TIMING=1 ./facade_roundtrip 3840 2160 1 0 11 8 2 12 1.0 6 2>&1 >/dev/null | grep "^facade-timing" | awk 'NR>2 {if ($2=="encode:") {et+=$4; er+=$10; ne++} else {dr+=$7; dt+=$10; nd++}} END {printf "encode: transform %.2f ms, run() %.2f ms (%d calls)\ndecode: run() %.2f ms, transform %.2f ms (%d calls)\n", et/ne, er/ne, ne, dr/nd, dt/nd, nd}'

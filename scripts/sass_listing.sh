#!/bin/bash
# SASS evidence that the contractions run on tcgen05 / TMEM / TMA: mnemonic counts per kernel + excerpts.
# usage: scripts/sass_listing.sh > profiles/sass_tcgen05.txt
LIB=radargnn_b200/lib/librgnn_b200.so
for K in node_gemm_kernel 'fused_layer_kernelILi0ELi2E'; do
  echo "==== $K (cuobjdump -sass $LIB)"
  cuobjdump -sass $LIB 2>/dev/null | awk -v k="$K" '/Function :/{on = index($0, k) > 0} on' > /tmp/sass_$$.txt
  echo "instructions: $(grep -c '^ *\/\*[0-9a-f]\{4,6\}\*\/' /tmp/sass_$$.txt)"
  for M in UTCHMMA UTCBAR UTMALDG UBLKCP UTCATOMSWS LDTM STTM SYNCS USETMAXREG FFMA2 FMNMX3 LDGSTS ATOMS STL LDL; do
    printf "  %-12s %s\n" $M "$(grep -c "[ .]$M" /tmp/sass_$$.txt)"
  done
  echo "-- first tensor-core / tensor-memory / bulk-copy instructions:"
  grep -m 12 -E "UTCHMMA|LDTM|STTM|UTMALDG|UBLKCP|UTCBAR|USETMAXREG" /tmp/sass_$$.txt | sed 's/  */ /g' | cut -c1-150
  rm -f /tmp/sass_$$.txt
done

#!/bin/bash
# K1 variants on one box (stand-alone bench, scripts/k1_bench.cu): timing + every-cell check against the scalar functor
# third sweep: ONE unit per CTA (ctas = 0 -> grid = number of units, dispatched in order by the hardware)
mkdir -p gpurun_out
out=gpurun_out/r02k3_k1_variants.log
: > $out
for b in build/k1v/k1_*; do
  echo "== $b" >> $out
  for tr in 1 2 4 8 16; do timeout 120 $b 85900 1 $tr 0 >> $out 2>&1; done
  for tr in 2 4 8; do
  timeout 120 $b 20000 0 $tr 0 >> $out 2>&1
  timeout 120 $b 13509 0 $tr 0 >> $out 2>&1
  timeout 120 $b 2392 1 $tr 0 >> $out 2>&1
  done
done
cut -c1-200 $out

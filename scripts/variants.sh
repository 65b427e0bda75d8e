#!/bin/bash
# Times every build variant under build/variants (scripts/build_variants.sh) on the GPU box,
# with the metric the library picks (norm-expanded integer scan) and with the dx/dy integer scan forced.
for f in "" build/variants/*.so; do
  for m in "" int; do
    echo -n "metric=${m:-auto} "
    TSPCUDA_FORCE_METRIC=$m TSPCUDA_LIB=${f:+$PWD/$f} timeout 180 python scripts/variant_time.py 2>&1 | tail -1
  done
done

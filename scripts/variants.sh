#!/bin/bash
# Runs $1 (a python script + args) under every build variant in build/variants.
for f in build/variants/*.so; do
  echo "== $f"
  TSPCUDA_LIB=$PWD/$f timeout 300 python "$@" 2>&1 | grep -v "auto \|full  "
done

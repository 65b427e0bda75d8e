#!/bin/bash
for f in build/variants/*.so; do TSPCUDA_LIB=$PWD/$f timeout 120 python scripts/variant_time.py 2>&1 | tail -1; done

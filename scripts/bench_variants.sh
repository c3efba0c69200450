#!/bin/bash
# usage: bench_variants.sh lib1.so lib2.so ... : E_loc kernel time with each library variant (env passed through)
for lib in "$@"; do
  cp "$lib" fermiflow_b200/libfermiflow_b200.so
  echo "== $lib"
  timeout 200 python scripts/dev_eloc_time.py 65536
done

#!/bin/bash
# usage: bench_variants2.sh "script args" lib1.so lib2.so ... : run a timing script with each library variant
cmd="$1"; shift
cp fermiflow_b200/libfermiflow_b200.so /tmp/lib_keep.so
for lib in "$@"; do
  cp "$lib" fermiflow_b200/libfermiflow_b200.so
  echo "== $lib"
  timeout 300 python $cmd 2>&1 | tail -2
done
cp /tmp/lib_keep.so fermiflow_b200/libfermiflow_b200.so

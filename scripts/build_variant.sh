#!/bin/bash
# build_variant.sh NAME [nvcc flags...]: libff_NAME.so with capi_eloc.cu compiled with the extra flags (dev A/B builds;
# scripts select it with FF_DEV_LIB=libff_NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC "$@" -c -o build/capi_eloc_$name.o fermiflow_b200/csrc/capi_eloc.cu 2>&1 | grep -v "177-D\|\^\|^$\|detected during\|Remark\|constexpr bool has_mu" || true
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o fermiflow_b200/libff_$name.so build/capi.o build/capi_eloc_$name.o

#!/bin/bash
# tools/build_variants.sh "TAG -DFOO=1 -DBAR=2" ... : build engine variants into fri-reinforcementlearning-c_b200/build/variants/friq_TAG.so
# (experiments only; select one at run time with FRIQ_LIB=<path>)
set -e
cd "$(dirname "$0")/../fri-reinforcementlearning-c_b200"
mkdir -p build/variants
rm -f build/variants/*.so
for cfg in "$@"; do
  set -- $cfg
  tag=$1; shift
  (
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -std=c++17 \
      "$@" -Xptxas -v -c csrc/friq_engine.cu -o build/variants/e_$tag.o 2> build/variants/ptxas_$tag.log &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/friq_$tag.so build/variants/e_$tag.o build/friq_host.o build/friq_multi.o build/frirl_dropin.o -ldl
    rm -f build/variants/e_$tag.o
  ) &
done
wait
ls build/variants/*.so

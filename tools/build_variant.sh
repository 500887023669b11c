#!/bin/bash
# usage: tools/build_variant.sh <name> <tu: w|tp|mdp|b200> "<extra nvcc flags>"  -> build/variants/libamdp_<name>.so
# (kernel A/B experiments: select at run time with AMDP_B200_LIB=build/variants/libamdp_<name>.so)
set -e
name=$1; tu=$2; flags=$3
cd "$(dirname "$0")/.."
mkdir -p build/variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude --fmad=false $flags \
  -Xptxas -v -c -o build/variants/amdp_${tu}_${name}.o xenoverse_b200/csrc/amdp_${tu}.cu 2> build/variants/${name}.ptxas.log
objs=""
for t in b200 mdp w tp vi sampler mt; do
  if [ $t = $tu ]; then objs="$objs build/variants/amdp_${tu}_${name}.o"; else objs="$objs build/amdp_$t.o"; fi
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libamdp_${name}.so $objs
grep -E "Used [0-9]+ reg" build/variants/${name}.ptxas.log | sort | uniq -c | head -5

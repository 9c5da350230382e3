#!/bin/bash
# Build K1s variants (separate libraries under build/, shipped to the GPU box by gpurun) -- run here, on the CPU box.
# usage: profiles/build_k1s_variants.sh name "-DSFM_KS_..." [name flags ...]
set -e -o pipefail
cd "$(dirname "$0")/.."
mkdir -p build
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags \
       -Xptxas -v -o build/libsfm_$name.so carla-social-force-model_b200/csrc/sfm_api.cu 2>&1 | grep -A3 "k1_sym_pairsILb0" | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores"| tr "\n" " " | sed "s/^/$name: /"
done

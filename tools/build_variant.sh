#!/bin/bash
# Developer aid: build/variants/<name>.so = the library with extra -D flags on poa_kernel.cu (A/B with tools/gpu_ab.sh).
# usage: tools/build_variant.sh <name> [-DFLAG ...]
set -euo pipefail
name=$1; shift
cd "$(dirname "$0")/../hypo_b200/csrc"
mkdir -p ../../build/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --std=c++17 -Xcompiler -fPIC,-Wall -ccbin /usr/bin/g++"
$NV "$@" -Xptxas -v -c -o /tmp/variant_$name.o poa_kernel.cu 2>&1 | grep -A1 "poa_kernelILb1ELb1ELb0ELi3ELi0" | grep -o "Used [0-9]* registers" || true
[ -f arms.o ] || make arms.o >/dev/null
[ -f support.o ] || make support.o >/dev/null
$NV "$@" -c -o /tmp/variant_api_$name.o api.cu   # (the arena layout is shared with the host side)
/usr/local/cuda/bin/nvcc -shared -o ../../build/variants/$name.so /tmp/variant_$name.o /tmp/variant_api_$name.o arms.o support.o -lcudart_static -lpthread -ldl -lrt 2>/dev/null
echo "built build/variants/$name.so"

#!/bin/bash
# scripts/build_variants.sh — A/B builds of libnsm_b200.so (CTA size / resident CTAs / b^-1 staging) into
# nimblesm_b200/lib/variants/; select one at run time with NSM_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../nimblesm_b200/csrc"
mkdir -p ../lib/variants
rm -f ../lib/variants/*
build() { # name, extra flags
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
     -Xcompiler -fPIC -ccbin /usr/bin/g++ $2 -shared -o ../lib/variants/libnsm_b200_$1.so nsm_b200.cu \
     -Xptxas -v 2> ../lib/variants/ptxas_$1.log &
}
build prefetch "-DNSM_BINV_PREFETCH"
build stage_elastic "-DNSM_BINV_STAGE_ELASTIC=1"
wait
for f in ../lib/variants/ptxas_*.log; do echo $f; grep -A2 "element_force_kernelILi1ELb0ELi2E\|element_force_kernelILi0ELb0ELi2E" $f | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores"; done

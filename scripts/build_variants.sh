#!/bin/bash
# scripts/build_variants.sh — A/B builds of libnsm_b200.so (CTA size / resident CTAs / ticket chunk) into
# nimblesm_b200/lib/variants/; select one at run time with NSM_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../nimblesm_b200/csrc"
mkdir -p ../lib/variants
build() { # name, extra flags
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
     -Xcompiler -fPIC -ccbin /usr/bin/g++ $2 -shared -o ../lib/variants/libnsm_b200_$1.so nsm_b200.cu \
     -Xptxas -v 2> ../lib/variants/ptxas_$1.log &
}
build chunk1 "-DNSM_TICKET_CHUNK=1"
build chunk4 "-DNSM_TICKET_CHUNK=4"
build chunk32 "-DNSM_TICKET_CHUNK=32"
build t128b4 "-DNSM_ELEM_THREADS=128 -DNSM_ELEM_MIN_BLOCKS=4"
wait
for f in ../lib/variants/ptxas_*.log; do echo $f; grep -A2 "element_force_kernelILi1ELb0ELi2E" $f | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores"; done

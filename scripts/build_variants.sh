#!/bin/bash
# scripts/build_variants.sh — A/B builds of libnsm_b200.so into nimblesm_b200/lib/variants/ (each with its own
# nsm_b200_kernel_info, so bench.py reports the variant's own instruction counts); select one at run time with
# NSM_B200_LIB=<path>.   usage: build_variants.sh name1="flags" name2="flags" ...
set -e
cd "$(dirname "$0")/../nimblesm_b200/csrc"
V=../lib/variants
mkdir -p $V
build() { # name, extra flags
  ( /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
      -Xcompiler -fPIC -ccbin /usr/bin/g++ $2 -c -o $V/nsm_b200_$1.o nsm_b200.cu -Xptxas -v 2> $V/ptxas_$1.log
    mkdir -p $V/inc_$1
    python ../../scripts/sass_hot_loop.py $V/nsm_b200_$1.o hex8_kernels.cuh hex8_math.cuh > $V/kernel_info_$1.json
    ( printf 'R"NSMJSON(' ; cat $V/kernel_info_$1.json ; printf ')NSMJSON"\n' ) > $V/inc_$1/kernel_info.inc
    /usr/bin/g++ -O2 -fPIC -std=c++17 -I$V/inc_$1 -c -o $V/kernel_info_$1.o kernel_info.cc
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $V/libnsm_b200_$1.so $V/nsm_b200_$1.o $V/kernel_info_$1.o
    rm -f $V/nsm_b200_$1.o $V/kernel_info_$1.o ) &
}
for spec in "$@"; do build "${spec%%=*}" "${spec#*=}"; done
wait
for f in $V/ptxas_*.log; do echo $f; grep -A2 "element_force_kernelILi1ELb0ELi2E\|element_force_kernelILi0ELb0ELi2E" $f | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores"; done

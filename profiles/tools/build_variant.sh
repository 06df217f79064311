#!/bin/bash
# usage: scratch/build_variant.sh <name> <source.cu> [-DFLAG=..]... : libbam_gpu_<name>.so with one translation unit rebuilt
name=$1; src=$2; shift 2
mkdir -p scratch/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin g++ -Xptxas -v "$@" -c bam_b200/csrc/$src -o scratch/variants/$name.o 2> scratch/variants/$name.ptxas.log || exit 1
objs=$(ls bam_b200/build/*.o | grep -v "/$src.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o scratch/variants/libbam_gpu_$name.so $objs scratch/variants/$name.o -lcudart -ldl

#!/bin/bash
# build_variant.sh <name> [nvcc flags...]: an alternative build of the library under build/libfsb_<name>.so (kernel experiments;
# load it with FSB_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../fundstababm.jl_b200/csrc"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC"
for f in fsb_abi fsb_step_big fsb_step_fixed fsb_facts; do nvcc $FLAGS "$@" -c -o /tmp/v_${name}_$f.o $f.cu & done; wait
mkdir -p ../../build
nvcc -shared -o ../../build/libfsb_$name.so /tmp/v_${name}_fsb_abi.o /tmp/v_${name}_fsb_step_big.o /tmp/v_${name}_fsb_step_fixed.o /tmp/v_${name}_fsb_facts.o -ldl
echo built build/libfsb_$name.so

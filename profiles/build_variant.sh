#!/bin/bash
# build_variant.sh NAME [-DSWITCH ...]: a build of the library with tuning switches, into
# build_variants/NAME.so (git-ignored; it travels to the GPU box).  Select it at run time with
# GOF_B200_LIB=build_variants/NAME.so (game_of_flies_b200/_lib.py).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC "$@" \
    game_of_flies_b200/csrc/gof_b200.cu -o build_variants/$name.so
echo build_variants/$name.so

#!/bin/bash
# Builds the clock64()-instrumented variant of the library (-DPCC_PROFILE) next to the product: gpurun_exp_prof/libpcc_b200_prof.so
# (git-ignored: gpurun_exp_*).  Use with PCC_B200_LIB=gpurun_exp_prof/libpcc_b200_prof.so python tools/phase_profile_packed.py
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_exp_prof
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -DPCC_PROFILE -Xcompiler -fPIC -shared \
     -Iinclude -o gpurun_exp_prof/libpcc_b200_prof.so pcc-rl_b200/csrc/pcc_b200.cu
echo built gpurun_exp_prof/libpcc_b200_prof.so

#!/bin/bash
# Rebuild libvolt_b200.so in-tree (same flags as __graft_entry__.build()).  Usage: tools/build.sh [-DVOLT_PROFILE ...]
set -e
cd "$(dirname "$0")/../volt_b200/csrc"
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v \
  -o libvolt_b200.so api.cu cov_build.cu chol_batched.cu chol_tc.cu chol_large.cu gp_predict.cu rollout.cu 2>&1 \
  | grep -E "error|warning|mll_batched_tc_kernel|rollout_kernel|large_" -A2 | grep -E "error|warning|Used|spill" || true
ls -la libvolt_b200.so

#!/bin/bash
# Rebuild the CUDA library in-tree (same flags as __graft_entry__.build()).
#   tools/build.sh             -> volt_b200/csrc/libvolt_b200.so
#   tools/build.sh --profile   -> volt_b200/csrc/libvolt_prof.so with the clock64() segment timers (tools/seg_probe.py)
set -e
cd "$(dirname "$0")/../volt_b200/csrc"
OUT=libvolt_b200.so; EXTRA=""
if [ "$1" = "--profile" ]; then OUT=libvolt_prof.so; EXTRA="-DVOLT_PROFILE"; fi
if [ "$1" = "--variant" ]; then OUT=libvolt_$2.so; EXTRA="$3"; fi   # experimental build: tools/build.sh --variant <name> "<-D flags>"
nvcc $EXTRA -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v \
  -o $OUT api.cu cov_build.cu chol_batched.cu chol_tc.cu chol_large.cu gp_predict.cu rollout.cu rollout_stats.cu gpcv.cu gemm_nt.cu 2>&1 \
  | grep -E "error|warning|mll_batched_tc_kernel|rollout_kernel|large_" -A2 | grep -E "error|warning|Used|spill" || true
ls -la $OUT

#!/bin/bash
# Builds the standalone micro-benchmarks of this directory into bin/ (sm_100a).
#   fp32_issue        issue rate of scalar vs packed FP32 instructions
#   stft_bench        k_stft alone: timing + double-precision spot check (STFT_FLAGS="-DAID_STFT_WARPS=2 ..." builds a variant)
#   stft_tc           the same harness on the experimental kernel of stft_tc.cuh (second transform on tcgen05); not yet run
#   umma_probe        one tcgen05 tile product (FP16 operands, no-swizzle K-major layout, FP32 accumulator in TMEM), plain and
#                     as the hi/lo split the STFT's second transform would use; checks descriptors and TMEM read-back
set -e
cd "$(dirname "$0")"
mkdir -p bin
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
S=../../audio_ident_b200/csrc/stft.cu
nvcc $F -o bin/fp32_issue fp32_issue.cu
nvcc $F $STFT_FLAGS -o bin/stft_bench stft_bench.cu $S
nvcc $F -o bin/umma_probe umma_probe.cu
nvcc $F -DAID_STFT_TC -o bin/stft_tc stft_bench.cu $S        # experimental: second transform on tcgen05 (stft_tc.cuh)
echo built

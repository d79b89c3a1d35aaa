#!/bin/bash
# A/B timing of alternative builds of libbgpt_cuda.so on one box: bash tools/ab_libs.sh <tag> "<lib suffixes, '-' = the default build>" "<formats>"
# per build and format: the decode loop (k_mega5<FMT, 0, 0>) at n_past 511 and the C++ sampling loop (k_mega5<FMT, 0, 1>) at n_past 384..639
TAG=$1; VARS=${2:--}; FTS=${3:-q4_0 q4_1 q5_0 q5_1 q8_0 f16}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in $VARS; do
  lib=""; [ "$v" != "-" ] && lib=$PWD/biogpt.cpp_b200/csrc/libbgpt_cuda_$v.so
  for ft in $FTS; do
    d=$(BGPT_CUDA_LIB=$lib timeout 120 python tools/profile_decode.py --ftype $ft --n-past 511 --steps 64 --warm 4 2>&1 | grep -o "[0-9.]* us/token")
    [ -z "$lib" ] && e=$(timeout 120 python tools/e2e_bench.py --ftype $ft --steps 192 --n-past 384 2>&1 | grep -o "C++ loop [0-9.]* us")
    echo "build=$v $ft decode-loop $d | sampler $e"
  done
done | tee $OUT/ab.log

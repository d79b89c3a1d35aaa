#!/bin/bash
# v4 bring-up: parity test of the generation-4 kernel, then phase stamps and timing
OUT=gpurun_out/${1:-v4}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -m gpu -k "generation4" > $OUT/pytest_v4.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_v4.log
for c in 147 0 40; do for np in 0 511; do BGPT_MEGA_PROF=1 BGPT_MEGA_PROF_CTA=$c timeout 300 python tools/profile_decode.py --n-past $np --steps 8 --warm 4; done; done > $OUT/phases.log 2>&1
cat $OUT/phases.log

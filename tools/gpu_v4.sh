#!/bin/bash
# v4 iteration: parity test of the generation-4 kernel, then the all-CTA trace and plain timing
OUT=gpurun_out/${1:-v4}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -m gpu -k "generation4" > $OUT/pytest_v4.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_v4.log
for np in 0 507; do timeout 200 python tools/trace_decode.py --n-past $np --dump $OUT/trace_$np.npz; done > $OUT/trace.log 2>&1
grep -v Warning $OUT/trace.log | grep -v "nanm\|return np"
for np in 0 511 1015; do timeout 200 python tools/profile_decode.py --n-past $np --steps 8 --warm 4; done 2>&1 | tee $OUT/timing.log

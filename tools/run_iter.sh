#!/bin/bash
# iteration visit: full GPU parity suite, then the per-operator schedule's benches (CUDA graph on / off) and the quantiser
OUT=gpurun_out/${1:-it}; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
for g in 1 0; do echo "== BGPT_GRAPH=$g"; BGPT_GRAPH=$g timeout 300 python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 64 | tail -1; BGPT_GRAPH=$g timeout 300 python tools/prompt_bench.py --ftype q8_0 --n 8,64; done 2>&1 | tee $OUT/graph.log
timeout 300 python tools/quant_bench.py 2>&1 | tee $OUT/quant.log

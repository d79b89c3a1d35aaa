#!/bin/bash
# multi-GPU visit (gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N>
TAG=$1; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests/test_replica_driver.py -m gpu -q > $OUT/pytest_replicas.log 2>&1; echo "replica tests rc=$?"; tail -2 $OUT/pytest_replicas.log
timeout 900 python bench.py --workload streams --driver cpp --gpus $N --steps 2 --warmup 1 > $OUT/bench_streams_cpp_${N}gpu.json 2> $OUT/bench_streams_cpp.err; echo "streams cpp rc=$?"; cut -c1-900 $OUT/bench_streams_cpp_${N}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --no-extras --cpu-budget 8 > $OUT/bench_decode_${N}gpu.json 2> $OUT/bench_decode.err; echo "decode torchrun rc=$?"; cut -c1-700 $OUT/bench_decode_${N}gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload streams --steps 2 --warmup 1 > $OUT/bench_streams_${N}gpu.json 2> $OUT/bench_streams.err; echo "streams torchrun rc=$?"; cut -c1-700 $OUT/bench_streams_${N}gpu.json

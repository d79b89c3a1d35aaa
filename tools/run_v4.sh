timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -m gpu -k "generation4" 2>&1 | tail -3
for np in 0 511 1015; do timeout 200 python tools/profile_decode.py --n-past $np --steps 8 --warm 4; done
BGPT_MEGA_PROF=1 timeout 200 python tools/trace_decode.py --n-past 507 2>&1 | grep -E "kernel|layer period|LN: sync1 -> before|P. x1* polled|P. all|publ|polled|done"

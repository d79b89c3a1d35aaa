#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + one full capture of k_mega.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [stages]   stages default "tests smoke bench launches full prompt"
TAG=${1:-r1}
STAGES=${2:-"tests smoke bench launches full prompt"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for st in $STAGES; do
case $st in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log ;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
  tail -3 $OUT/smoke.log ;;
bench)
  timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
  cat $OUT/bench.json ;;
benchref)
  timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "benchref rc=$?"
  cat $OUT/bench_ref.json ;;
phases)
  for np in 0 511 1023; do BGPT_MEGA_PROF=1 timeout 300 python tools/profile_decode.py --n-past $np --steps 8 --warm 4; done > $OUT/phases.log 2>&1
  cat $OUT/phases.log ;;
sweep)
  for ft in f16 q4_0 q4_1 q5_0 q5_1 q8_0; do timeout 300 python tools/profile_decode.py --ftype $ft --n-past 511 --steps 32 --warm 8 | head -1; done > $OUT/sweep.log 2>&1
  cat $OUT/sweep.log ;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_decode.csv \
      python tools/profile_decode.py --n-past 511 --steps 8 --warm 0 > $OUT/launches_decode.log 2>&1; echo "launches rc=$?" ;;
full)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mega -s 3 -c 1 -f -o $OUT/mega_full \
      python tools/profile_decode.py --n-past 511 --steps 5 --warm 0 > $OUT/mega_full.log 2>&1; echo "full rc=$?"
  ncu -i $OUT/mega_full.ncu-rep --page raw --csv > $OUT/mega_full_raw.csv 2>/dev/null
  tail -3 $OUT/mega_full.log ;;
promptfull)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc_q -s 8 -c 1 -f -o $OUT/gemm_tc_full \
      python tools/profile_prompt.py --ftype q8_0 --n 1024 > $OUT/gemm_tc_full.log 2>&1; echo "promptfull rc=$?"
  ncu -i $OUT/gemm_tc_full.ncu-rep --page raw --csv > $OUT/gemm_tc_full_raw.csv 2>/dev/null
  tail -3 $OUT/gemm_tc_full.log ;;
trace)
  for np in 0 507; do BGPT_MEGA_PROF=1 timeout 200 python tools/trace_decode.py --n-past $np; done > $OUT/trace.log 2>&1; grep -v "Warning\|nanm\|return np" $OUT/trace.log | head -60 ;;
streams)
  timeout 600 python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 64 > $OUT/streams.log 2>&1; cat $OUT/streams.log ;;
prompt)
  timeout 600 python tools/prompt_bench.py --ftype q8_0 --n 8,64,256,1024 > $OUT/prompt.log 2>&1; cat $OUT/prompt.log ;;
promptncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_prompt.csv \
      python tools/profile_prompt.py --ftype q8_0 --n 1024 > $OUT/launches_prompt.log 2>&1; echo "promptncu rc=$?" ;;
esac
done

#!/bin/bash
# GPU-box visit for the fused skinny-batch schedule: bash tools/gpu_sk.sh <tag> "<stages>"
TAG=${1:-sk}
STAGES=${2:-"check tests perf"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for st in $STAGES; do
case $st in
check)
  timeout 600 python tools/skinny_check.py > $OUT/check.log 2>&1; echo "check rc=$?"; cat $OUT/check.log | tail -40 ;;
tests)
  timeout 1500 python -m pytest tests/test_gpu_eval.py -m gpu -x -q -k "skinny or base_shape or generation4" > $OUT/pytest_sk.log 2>&1; echo "pytest rc=$?"
  tail -15 $OUT/pytest_sk.log ;;
alltests)
  timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
  tail -20 $OUT/pytest_gpu.log ;;
perf)
  for v in "X=0" "BGPT_SK_FC1_NW=32" "BGPT_SK_TN_PROJ=8" "BGPT_SK_TN_PROJ=8 BGPT_SK_FC1_NW=32" "X=1"; do
    echo "== variant: ${v:-default}"
    env $v timeout 300 python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 64
    env $v timeout 300 python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 32 --n-past 480 --reps 3
    env $v timeout 300 python tools/prompt_bench.py --ftype q8_0 --n 8
  done > $OUT/perf.log 2>&1
  cat $OUT/perf.log ;;
skips)
  # in-graph cost of each kernel kind: the step time with that kind of launch left out (results are garbage, timing only)
  for v in "X=0"; do
  for k in 0 1 2 4 8 16 32 64; do
    echo "== $v BGPT_SK_SKIP=$k (1 qkv, 2 attention, 4 out_proj, 8 fc1, 16 fc2, 32 lm_head, 64 LayerNorm kernels)"
    env $v BGPT_SK_SKIP=$k timeout 300 python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 32 --n-past 480 --reps 2
  done; done > $OUT/skips.log 2>&1
  cat $OUT/skips.log ;;
nograph)
  BGPT_GRAPH=0 timeout 600 python tools/skinny_check.py --ftypes q4_0,q5_1 > $OUT/check_nograph.log 2>&1; echo "nograph rc=$?"; tail -2 $OUT/check_nograph.log
  BGPT_PDL=0 timeout 600 python tools/skinny_check.py --ftypes q4_1,q8_0 > $OUT/check_nopdl.log 2>&1; echo "nopdl rc=$?"; tail -2 $OUT/check_nopdl.log ;;
nstreams)
  for S in 2 4 16 24; do timeout 300 python tools/streams_bench.py --ftype q5_1 --streams $S --steps 64; timeout 300 python tools/streams_bench.py --ftype q5_1 --streams $S --steps 32 --n-past 480; done > $OUT/nstreams.log 2>&1
  for ft in q4_0 q8_0; do timeout 300 python tools/streams_bench.py --ftype $ft --streams 8 --steps 64; done >> $OUT/nstreams.log 2>&1
  cat $OUT/nstreams.log ;;
launches)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_streams.csv \
      python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 2 --n-past 511 --reps 1 > $OUT/launches_streams.log 2>&1; echo "launches rc=$?"
  python tools/launch_summary.py $OUT/launches_streams.csv | tee $OUT/launches_streams_summary.txt
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_sk|k_embed|k_act|k_gemv" -c 400 --csv --log-file $OUT/launches_prompt8.csv \
      python tools/profile_prompt.py --ftype q8_0 --n 8 --n-past 512 --reps 2 > $OUT/launches_prompt8.log 2>&1; echo "launches prompt rc=$?"
  python tools/launch_summary.py $OUT/launches_prompt8.csv | tee $OUT/launches_prompt8_summary.txt ;;
full)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sk_mm -s 6 -c 5 -f -o $OUT/sk_mm_full \
      python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 2 --n-past 511 --reps 1 > $OUT/sk_mm_full.log 2>&1; echo "full rc=$?"
  ncu -i $OUT/sk_mm_full.ncu-rep --page raw --csv > $OUT/sk_mm_full_raw.csv 2>/dev/null
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sk_attn -s 1 -c 1 -f -o $OUT/sk_attn_full \
      python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 2 --n-past 511 --reps 1 > $OUT/sk_attn_full.log 2>&1; echo "full attn rc=$?"
  ncu -i $OUT/sk_attn_full.ncu-rep --page raw --csv > $OUT/sk_attn_full_raw.csv 2>/dev/null
  rm -f $OUT/*.ncu-rep.tmp; ls -la $OUT ;;
bench)
  timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
  cat $OUT/bench.json; tail -3 $OUT/bench.err ;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log ;;
esac
done

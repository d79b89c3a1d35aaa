#!/bin/bash
# round-2 GPU visits: bash tools/gpu_r2.sh <tag> <stages...>
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
for st in "$@"; do
case $st in
probe)
  for t in T1 T2 T5 T3 T4; do timeout 40 tools/_build/cluster_probe $t >> $OUT/cluster_probe.log 2>&1; echo "probe $t rc=$?" >> $OUT/cluster_probe.log; done; cat $OUT/cluster_probe.log ;;
tests)
  timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  grep -E "passed|failed|FAILED|Error|max\|d|tcgen05 prompt" $OUT/pytest_gpu.log | tail -40 ;;
newtests)
  timeout 1800 python -m pytest tests/test_gpu_eval.py -m gpu -q --maxfail=12 -s -k "tensor_core or base_model or eval_path" > $OUT/pytest_new.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_new.log
  grep -E "passed|failed|FAILED|Error|max\|d|tcgen05 prompt" $OUT/pytest_new.log | tail -40 ;;
gen5)
  timeout 900 python -m pytest tests/test_gpu_eval.py -m gpu -q --maxfail=3 -x -k "persistent_generations" > $OUT/pytest_gen5.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gen5.log
  grep -E "passed|failed|FAILED|Error|differ|timed out" $OUT/pytest_gen5.log | tail -20 ;;
headline)
  timeout 900 python -m pytest tests/test_gpu_eval.py -m gpu -q --maxfail=3 -x -k "base_model_headline or base_model_64" > $OUT/pytest_headline.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_headline.log
  grep -E "passed|failed|FAILED|Error|differ|timed out" $OUT/pytest_headline.log | tail -20 ;;
tcx)
  timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_eval.py -m gpu -q --maxfail=6 -k "tensor_core_exact or large_prompt or eval_path_map or opt_in" > $OUT/pytest_tcx.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_tcx.log
  grep -E "passed|failed|FAILED|Error|differ|timed out" $OUT/pytest_tcx.log | tail -20 ;;
tcxkinds)
  for kind in f16 i8; do
    BGPT_TCX_KIND=$kind timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_eval.py -m gpu -q --maxfail=6 -k "tensor_core_exact or large_prompt" > $OUT/pytest_tcx_$kind.log 2>&1; echo "pytest $kind rc=$?" | tee -a $OUT/pytest_tcx_$kind.log
    grep -E "passed|failed|FAILED|Error|differ" $OUT/pytest_tcx_$kind.log | tail -8
    for ft in q8_0 q4_0; do BGPT_TCX_KIND=$kind timeout 600 python tools/prompt_bench.py --ftype $ft --n 128,1024 > $OUT/prompt_${kind}_$ft.log 2>&1; cat $OUT/prompt_${kind}_$ft.log; done
  done
  BGPT_TCX_MIN_ROWS=32 timeout 600 python tools/prompt_bench.py --ftype q8_0 --n 32,64,96 > $OUT/prompt_tcx_small_q8_0.log 2>&1; cat $OUT/prompt_tcx_small_q8_0.log
  timeout 600 python tools/prompt_bench.py --ftype q8_0 --n 32,64,96 > $OUT/prompt_sk_small_q8_0.log 2>&1; cat $OUT/prompt_sk_small_q8_0.log ;;
promptsweep)
  for ft in ${FTYPES:-q8_0}; do
    BGPT_TCX_MIN_ROWS=0 timeout 600 python tools/prompt_bench.py --ftype $ft --n 128,256,1024 > $OUT/prompt_skinny_$ft.log 2>&1; cat $OUT/prompt_skinny_$ft.log
    timeout 600 python tools/prompt_bench.py --ftype $ft --n 8,64,128,256,1024 > $OUT/prompt_tcx_$ft.log 2>&1; cat $OUT/prompt_tcx_$ft.log
  done ;;
sanitize)
  # compute-sanitizer on the tiny / small models through the C ABI: memcheck over prompt + decode on every schedule, racecheck on the decode kernel
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/sanitizer_memcheck.log
  tail -4 $OUT/sanitizer_memcheck.log
  BGPT_M5_WATCHDOG_MCYC=400000 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py --quick > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/sanitizer_racecheck.log
  tail -4 $OUT/sanitizer_racecheck.log ;;
hostlib)
  timeout 600 python -m pytest tests/test_host_lib.py tests/test_replica_driver.py -m gpu -q --maxfail=6 > $OUT/pytest_host.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_host.log
  grep -E "passed|failed|FAILED|Error|differ" $OUT/pytest_host.log | tail ;;
ncu5)
  BGPT_M5_COOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_decode.csv \
      python tools/profile_decode.py --n-past 511 --steps 8 --warm 0 > $OUT/launches_decode.log 2>&1; echo "launches rc=$?"
  BGPT_M5_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mega5 -s 3 -c 1 -f -o $OUT/mega5_full \
      python tools/profile_decode.py --n-past 511 --steps 5 --warm 0 > $OUT/mega5_full.log 2>&1; echo "full rc=$?"
  ncu -i $OUT/mega5_full.ncu-rep --page raw --csv > $OUT/mega5_full_raw.csv 2>/dev/null
  tail -3 $OUT/mega5_full.log ;;
ncutcx)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_prompt.csv \
      python tools/profile_prompt.py --ftype q8_0 --n 1024 > $OUT/launches_prompt.log 2>&1; echo "promptncu rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc_x -s 8 -c 1 -f -o $OUT/gemm_tcx_full \
      python tools/profile_prompt.py --ftype q8_0 --n 1024 > $OUT/gemm_tcx_full.log 2>&1; echo "tcxfull rc=$?"
  ncu -i $OUT/gemm_tcx_full.ncu-rep --page raw --csv > $OUT/gemm_tcx_full_raw.csv 2>/dev/null
  tail -3 $OUT/gemm_tcx_full.log ;;
e2esweep)
  for ft in q4_0 q4_1 q5_0 q5_1 q8_0 f16; do for cfg in "1 1" "0 1" "0 0"; do set -- $cfg; echo "$ft chain=$1 tail=$2"; BGPT_CHAIN=$1 BGPT_TOPK_TAIL=$2 timeout 120 python tools/e2e_bench.py --ftype $ft --steps 256 --n-past 384 2>&1 | grep "C++ loop"; done; done | tee $OUT/e2esweep.log ;;
ncutk)
  # the sampler's instantiation of the decode kernel (k_mega5<.., TK>): launch list of the C++ sampling loop, then --set full of one launch
  BGPT_CHAIN=0 BGPT_M5_COOP=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 60 --csv --log-file $OUT/launches_e2e.csv \
      python tools/e2e_bench.py --ftype q4_0 --steps 8 --n-past 511 > $OUT/launches_e2e.log 2>&1; echo "launches rc=$?"
  BGPT_CHAIN=0 BGPT_M5_COOP=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:ELb0ELb1EEv8M5Params -s 6 -c 1 -f -o $OUT/mega5tk_full \
      python tools/e2e_bench.py --ftype q4_0 --steps 8 --n-past 511 > $OUT/mega5tk_full.log 2>&1; echo "full rc=$?"
  ncu -i $OUT/mega5tk_full.ncu-rep --page raw --csv > $OUT/mega5tk_full_raw.csv 2>/dev/null
  rm -f $OUT/mega5tk_full.ncu-rep
  tail -3 $OUT/mega5tk_full.log; head -5 $OUT/launches_e2e.csv | cut -c1-300 ;;
fmtsweep)
  for ft in q4_0 q4_1 q5_0 q5_1 q8_0 f16; do timeout 120 python tools/profile_decode.py --ftype $ft --n-past 511 --steps 64 --warm 4 2>&1 | grep "us/token"; done | tee $OUT/fmtsweep.log ;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
  tail -5 $OUT/smoke.log ;;
bench)
  timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
  cat $OUT/bench.json; tail -3 $OUT/bench.err ;;
trace5)
  for np in 4 507 1015; do BGPT_MEGA_PROF=1 timeout 200 python tools/trace_decode5.py --n-past $np; done > $OUT/trace.log 2>&1; grep -v "Warning\|nanm\|return np" $OUT/trace.log | head -150 ;;
rows)
  timeout 1500 python -m pytest tests/test_gpu_eval.py -m gpu -q --maxfail=4 -x -k "multi_row or eval_path_map or skinny" > $OUT/pytest_rows.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_rows.log
  grep -E "passed|failed|FAILED|Error|differ|timed out" $OUT/pytest_rows.log | tail -20 ;;
rowsbench)
  for bp in 2 1; do
    BGPT_BATCH_PATH=$bp timeout 300 python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 64 --n-past 0 2>&1 | tail -1
    BGPT_BATCH_PATH=$bp timeout 300 python tools/streams_bench.py --ftype q5_1 --streams 8 --steps 64 --n-past 900 2>&1 | tail -1
    BGPT_BATCH_PATH=$bp timeout 300 python tools/streams_bench.py --ftype q5_1 --streams 4 --steps 64 --n-past 448 2>&1 | tail -1
    BGPT_BATCH_PATH=$bp timeout 300 python tools/streams_bench.py --ftype q4_0 --streams 2 --steps 64 --n-past 448 2>&1 | tail -1
    BGPT_BATCH_PATH=$bp BGPT_TCX_MIN_ROWS=0 timeout 300 python tools/prompt_bench.py --ftype q8_0 --n 8,4,2 2>&1 | tail -3
  done > $OUT/rowsbench.log 2>&1
  cat $OUT/rowsbench.log ;;
rowstrace)
  for np in 8 511 1000; do BGPT_MEGA_PROF=1 timeout 300 python tools/trace_rows.py --ftype q5_1 --rows 8 --mode streams --n-past $np; done > $OUT/rows_trace.log 2>&1
  BGPT_MEGA_PROF=1 timeout 300 python tools/trace_rows.py --ftype q8_0 --rows 8 --mode prompt --n-past 504 >> $OUT/rows_trace.log 2>&1
  cat $OUT/rows_trace.log ;;
tcw)
  timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_eval.py -m gpu -q --maxfail=8 -s -k "tma_fed or tensor_core_f16 or large_prompt or eval_path_map or f16_prompt" > $OUT/pytest_tcw.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_tcw.log
  grep -E "passed|failed|FAILED|Error|differ|timed out|max\|d" $OUT/pytest_tcw.log | tail -40 ;;
tcwbench)
  for ft in ${FTYPES:-q8_0 q4_0 f16}; do BGPT_F16_TC_MIN_ROWS=32 timeout 600 python tools/prompt_bench.py --ftype $ft --n 32,128,256,1024 2>&1 | tail -4; done > $OUT/prompt_tcw.log 2>&1
  BGPT_TCW=0 timeout 600 python tools/prompt_bench.py --ftype q8_0 --n 128,1024 2>&1 | tail -2 >> $OUT/prompt_tcw.log
  BGPT_F16_TC_MIN_ROWS=0 timeout 600 python tools/prompt_bench.py --ftype f16 --n 32,128,1024 2>&1 | tail -3 >> $OUT/prompt_tcw.log
  cat $OUT/prompt_tcw.log ;;
ncutcw)
  timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -s -k "tensor_core_f16" 2>&1 | grep -E "max\|d|passed|failed" > $OUT/f16_mm_err.log; cat $OUT/f16_mm_err.log
  for ft in ${FTYPES:-q8_0 f16}; do
    BGPT_F16_TC_MIN_ROWS=32 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_prompt_$ft.csv \
        python tools/profile_prompt.py --ftype $ft --n 1024 > $OUT/launches_prompt_$ft.log 2>&1; echo "promptncu $ft rc=$?"
    python tools/launch_summary.py $OUT/launches_prompt_$ft.csv 2>&1 | tail -12
  done ;;
ncutcwfull)
  BGPT_F16_TC_MIN_ROWS=32 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tcw_f16 -s 4 -c 4 -f -o $OUT/tcw_f16_full \
      python tools/profile_prompt.py --ftype f16 --n 1024 > $OUT/tcw_f16_full.log 2>&1; echo "f16 full rc=$?"
  ncu -i $OUT/tcw_f16_full.ncu-rep --page raw --csv > $OUT/tcw_f16_full_raw.csv 2>/dev/null
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tcw_exact -s 4 -c 4 -f -o $OUT/tcw_exact_full \
      python tools/profile_prompt.py --ftype q8_0 --n 1024 > $OUT/tcw_exact_full.log 2>&1; echo "exact full rc=$?"
  ncu -i $OUT/tcw_exact_full.ncu-rep --page raw --csv > $OUT/tcw_exact_full_raw.csv 2>/dev/null
  ls -la $OUT ;;
tcwab)
  for ft in q8_0 q5_1; do
    for tpt in 4 8; do echo "BGPT_TCW_TPT=$tpt"; BGPT_TCW_TPT=$tpt timeout 600 python tools/prompt_bench.py --ftype $ft --n 128,1024 2>&1 | tail -2; done
  done > $OUT/prompt_tcw_ab.log 2>&1
  for sp in 1 0; do echo "BGPT_F16_TC_SPLIT=$sp"; BGPT_F16_TC_MIN_ROWS=32 BGPT_F16_TC_SPLIT=$sp timeout 600 python tools/prompt_bench.py --ftype f16 --n 32,128,1024 2>&1 | tail -3
    BGPT_F16_TC_SPLIT=$sp timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_eval.py -m gpu -q -s -k "tensor_core_f16 or f16_prompt" 2>&1 | grep -E "max\|d|passed|failed"; done >> $OUT/prompt_tcw_ab.log 2>&1
  cat $OUT/prompt_tcw_ab.log ;;
e2e)
  timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_host_lib.py -m gpu -q --maxfail=5 -k "topk or sampler or chained" > $OUT/pytest_topk.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_topk.log
  grep -E "passed|failed|FAILED|Error|assert" $OUT/pytest_topk.log | tail -12
  timeout 300 python tools/e2e_bench.py --ftype q4_0 --steps 256 > $OUT/e2e.log 2>&1
  BGPT_CHAIN=0 timeout 300 python tools/e2e_bench.py --ftype q4_0 --steps 256 >> $OUT/e2e.log 2>&1
  BGPT_CHAIN=0 BGPT_TOPK_TAIL=0 timeout 300 python tools/e2e_bench.py --ftype q4_0 --steps 256 >> $OUT/e2e.log 2>&1
  timeout 300 python tools/e2e_bench.py --ftype f16 --steps 256 >> $OUT/e2e.log 2>&1
  cat $OUT/e2e.log ;;
f16g5)
  timeout 1500 python -m pytest tests/test_gpu_eval.py -m gpu -q --maxfail=3 -x -k "persistent_generations and f16 or base_shape_matches_oracle and f16 or base_model_64 and f16" > $OUT/pytest_f16g5.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_f16g5.log
  grep -E "passed|failed|FAILED|Error|differ|timed out|generation" $OUT/pytest_f16g5.log | tail -20
  for np in 0 511 980; do timeout 300 python tools/profile_decode.py --ftype f16 --n-past $np --steps 32 --warm 8 | head -1; done > $OUT/decode_f16.log 2>&1
  cat $OUT/decode_f16.log ;;
attn2)
  timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_eval.py -m gpu -q --maxfail=6 -k "attention or large_prompt or small_matches or tiny_matches or f16_prompt" > $OUT/pytest_attn2.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_attn2.log
  grep -E "passed|failed|FAILED|Error|differ" $OUT/pytest_attn2.log | tail -12
  for v in 2 1; do echo "BGPT_ATTN_TILE=$v"; BGPT_ATTN_TILE=$v timeout 600 python tools/prompt_bench.py --ftype q8_0 --n 128,1024 2>&1 | tail -2
    BGPT_ATTN_TILE=$v BGPT_F16_TC_MIN_ROWS=32 timeout 600 python tools/prompt_bench.py --ftype f16 --n 8,1024 2>&1 | tail -2; done > $OUT/prompt_attn2.log 2>&1
  cat $OUT/prompt_attn2.log ;;
decode)
  for ft in ${FTYPES:-q4_0}; do for np in 0 511 980; do timeout 300 python tools/profile_decode.py --ftype $ft --n-past $np --steps 32 --warm 8 | head -1; done; done > $OUT/decode.log 2>&1
  cat $OUT/decode.log ;;
esac
done

#!/bin/bash
# Round-2 GPU session helper: bash scripts/r02_run.sh TAG "sections"
set -u
TAG=${1:-r02a}
SECTIONS=${2:-"mk slmft bench"}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $SECTIONS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if has mk; then
  timeout 900 python -m pytest tests/test_decode_mk_gpu.py -x -q > $OUT/${TAG}_mk.log 2>&1; echo "exit $?" >> $OUT/${TAG}_mk.log
  tail -25 $OUT/${TAG}_mk.log
fi
if has slmft; then
  timeout 1200 python -m pytest tests/test_slmft_gpu.py -x -q > $OUT/${TAG}_slmft.log 2>&1; echo "exit $?" >> $OUT/${TAG}_slmft.log
  tail -15 $OUT/${TAG}_slmft.log
fi
if has alltests; then
  timeout 1800 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "exit $?" >> $OUT/${TAG}_tests.log
  tail -15 $OUT/${TAG}_tests.log
fi
if has bench; then
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err
  tail -c 3000 $OUT/${TAG}_bench_bf16.json; tail -5 $OUT/${TAG}_bench_bf16.err
fi
if has trace; then
  timeout 600 python scripts/decode_trace.py > $OUT/${TAG}_trace.txt 2>&1
  cat $OUT/${TAG}_trace.txt
fi
if has b1; then
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-leg --workload vico_b1 > $OUT/${TAG}_bench_b1.json 2> $OUT/${TAG}_bench_b1.err
  tail -c 1500 $OUT/${TAG}_bench_b1.json; tail -5 $OUT/${TAG}_bench_b1.err
fi
if has trace1; then
  timeout 600 python scripts/decode_trace.py 1 300 > $OUT/${TAG}_trace_b1.txt 2>&1
  grep -v "  phase" $OUT/${TAG}_trace_b1.txt
fi
if has barab; then
  for m in 0 1; do
    DIM_MK_NOPS=8 DIM_MK_BAR=$m timeout 300 python scripts/decode_trace.py 256 300 bf16 > $OUT/${TAG}_barab_$m.txt 2>&1
    echo "bar mode $m"; grep -v "  phase" $OUT/${TAG}_barab_$m.txt; grep "nop" $OUT/${TAG}_barab_$m.txt | head -3
  done
fi
if has stages2; then
  DIM_MK_ATTN_STAGES=2 timeout 300 python scripts/decode_trace.py 256 300 bf16 > $OUT/${TAG}_stages2.txt 2>&1
  echo "2-stage ring"; grep -v "  phase" $OUT/${TAG}_stages2.txt
fi
if has attnffma; then
  DIM_MK_ATTN_FFMA=1 timeout 300 python scripts/decode_trace.py 256 300 bf16 > $OUT/${TAG}_attnffma.txt 2>&1
  echo "FFMA attention items"; grep -v "  phase" $OUT/${TAG}_attnffma.txt
fi
if has ncumk; then
  # one --set full capture of the persistent decode kernel (bf16, 256 x 300): where do the warps stall?
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:decode_megakernel -c 1 -o $OUT/${TAG}_ncu_mk -f \
      python scripts/decode_trace.py 256 ${NCU_T:-300} bf16 > $OUT/${TAG}_ncu_mk.log 2>&1
  ncu -i $OUT/${TAG}_ncu_mk.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_mk.raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_ncu_mk.ncu-rep --page source --csv --print-source cuda,sass > $OUT/${TAG}_ncu_mk.source.csv 2>/dev/null
  ncu -i $OUT/${TAG}_ncu_mk.ncu-rep --page details > $OUT/${TAG}_ncu_mk.details.txt 2>/dev/null
  ls -la $OUT/${TAG}_ncu_mk*; tail -3 $OUT/${TAG}_ncu_mk.log
  gzip -f $OUT/${TAG}_ncu_mk.source.csv
fi
if has chains1; then
  DIM_MK_CHAINS=1 timeout 300 python scripts/decode_trace.py 256 300 bf16 > $OUT/${TAG}_chains1.txt 2>&1
  echo "single chain"; grep -v "  phase" $OUT/${TAG}_chains1.txt
  DIM_MK_STAGGER=0 timeout 300 python scripts/decode_trace.py 256 300 bf16 > $OUT/${TAG}_stagger0.txt 2>&1
  echo "two chains, no stagger"; grep -v "  phase" $OUT/${TAG}_stagger0.txt
  DIM_MK_STAGGER=5 timeout 300 python scripts/decode_trace.py 256 300 bf16 > $OUT/${TAG}_stagger5.txt 2>&1
  echo "two chains, stagger 5"; grep -v "  phase" $OUT/${TAG}_stagger5.txt
fi
if has mkdebug; then
  DIM_MK_DEBUG=1 timeout 300 python scripts/decode_trace.py 256 60 bf16 2>&1 | grep -v "  phase" | head -20
fi
if has mkcheck; then
  for cfg in "2 0" "2 2" "3 0"; do
    set -- $cfg
    echo "== chains=$1 stagger=$2"
    DIM_MK_CHAINS=$1 DIM_MK_STAGGER=$2 timeout 300 python scripts/mk_check.py 130 10 fp32_tc 2>&1 | tail -3
    DIM_MK_CHAINS=$1 DIM_MK_STAGGER=$2 timeout 300 python scripts/mk_check.py 256 8 fp32_tc 2>&1 | tail -3
  done
fi
if has chainsab; then
  for cfg in "2 0" "2 2" "2 5" "2 12" "3 0" "1 0"; do
    set -- $cfg
    echo "== chains=$1 stagger=$2"
    DIM_MK_CHAINS=$1 DIM_MK_STAGGER=$2 timeout 300 python scripts/decode_trace.py 256 300 bf16 2>&1 | grep -v "  phase"
  done
fi
if has smoke; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "exit $?" >> $OUT/${TAG}_smoke.log; tail -3 $OUT/${TAG}_smoke.log
fi
if has benchfull; then
  timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err
  tail -c 600 $OUT/${TAG}_bench_bf16.err
  python - $OUT/${TAG}_bench_bf16.json <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value %.0f ms %.1f e2e %.0f launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
r = d["roofline"]; print("roofline", r["kernel"], "frac %.3f" % r["frac"], "attn phases", r.get("attention_phases", {}).get("frac"), "share", r["share_of_step"])
print("phases", r.get("phases"))
p = d.get("fp32_parity_mode"); print("parity", p and (p["value"], p["ms_per_step"], p["e2e"]["value"], p["roofline"]["frac"]))
print("vq", d.get("vq_lookup")); print("others", d.get("other_workloads")); print("cpu", d.get("cpu_baseline")); print("clocks", d["clocks"])
for k in d["kernels"][:8]: print("   ", k)
PY
fi
if has benchref; then
  timeout 900 python bench.py --steps 3 --warmup 1 --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
  cut -c1-700 $OUT/${TAG}_bench_reference.json
fi
if has bf16contract; then
  timeout 900 python -m pytest tests/test_bf16_contract_gpu.py tests/test_decode_mk_gpu.py -q -s > $OUT/${TAG}_bf16contract.log 2>&1; echo "exit $?" >> $OUT/${TAG}_bf16contract.log
  grep -E "bf16|passed|failed|Error|assert" $OUT/${TAG}_bf16contract.log | head -30
fi
if has argmintc; then
  timeout 900 python -m pytest tests/test_vq_argmin_tc_gpu.py tests/test_ops_gpu.py tests/test_vqvae_gpu.py -q -s -x > $OUT/${TAG}_argmintc.log 2>&1; echo "exit $?" >> $OUT/${TAG}_argmintc.log
  grep -E "tokens with|passed|failed|Error|assert|exit" $OUT/${TAG}_argmintc.log | head -20
  timeout 300 python scripts/vq_roofline.py > $OUT/${TAG}_vq_roofline.jsonl 2> $OUT/${TAG}_vq_roofline.err
  grep argmin $OUT/${TAG}_vq_roofline.jsonl | cut -c1-300; tail -2 $OUT/${TAG}_vq_roofline.err
fi
if has argmindbg; then
  VQ_TC_DBG=${VQ_TC_DBG:-0,1,2,4,8,9,11,15} timeout 300 python scripts/vq_roofline.py 2>&1 | grep ablation
fi
if has argmintrace; then
  timeout 300 python scripts/vq_tc_trace.py 2>&1 | tail -30
fi
if has argmintrace2; then
  for d in ${VQ_TC_DBGS:-9 4 13}; do echo "== dbg $d"; VQ_TC_DBG=$d timeout 300 python scripts/vq_tc_trace.py 2>&1 | tail -20 | grep "tile  [89]"; done
fi
if has ncuargmin; then
  VQ_NCU=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:vq_argmin_tc -c 1 -o $OUT/${TAG}_ncu_argmin -f \
      python scripts/vq_roofline.py > $OUT/${TAG}_ncu_argmin.log 2>&1
  ncu -i $OUT/${TAG}_ncu_argmin.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_argmin.raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_ncu_argmin.ncu-rep --page details > $OUT/${TAG}_ncu_argmin.details.txt 2>/dev/null
  ncu -i $OUT/${TAG}_ncu_argmin.ncu-rep --page source --csv --print-source sass > $OUT/${TAG}_ncu_argmin.source.csv 2>/dev/null
  gzip -f $OUT/${TAG}_ncu_argmin.source.csv
  ls -la $OUT/${TAG}_ncu_argmin*; tail -3 $OUT/${TAG}_ncu_argmin.log
fi
if has mkab; then
  echo "== default"; timeout 300 python scripts/decode_trace.py 256 300 bf16 2>&1 | grep -v "  phase"
  echo "== 2-stage ring"; DIM_MK_ATTN_STAGES=2 timeout 300 python scripts/decode_trace.py 256 300 bf16 2>&1 | grep -v "  phase"
  echo "== no gelu fuse"; DIM_MK_NO_GELU_FUSE=1 timeout 300 python scripts/decode_trace.py 256 300 bf16 2>&1 | grep -v "  phase"
fi
if has speaker; then
  timeout 900 python -m pytest tests/test_vqspeaker_gpu.py -x -q > $OUT/${TAG}_speaker.log 2>&1; echo "exit $?" >> $OUT/${TAG}_speaker.log
  tail -15 $OUT/${TAG}_speaker.log
fi
if has slm; then
  timeout 900 python -m pytest tests/test_slm_gpu.py tests/test_compat_gpu.py -x -q > $OUT/${TAG}_slm.log 2>&1; echo "exit $?" >> $OUT/${TAG}_slm.log
  tail -25 $OUT/${TAG}_slm.log
fi
if has loader; then
  timeout 900 python -m pytest tests/test_loader_gpu.py -x -q > $OUT/${TAG}_loader.log 2>&1; echo "exit $?" >> $OUT/${TAG}_loader.log
  tail -25 $OUT/${TAG}_loader.log
fi
if has launchlist; then
  # launch list of the bench command (ncu replays every launch: serialised, cold-cache times -- shares, not absolutes)
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-leg --no-other-workloads > $OUT/${TAG}_launches_bench.log 2>&1
  python scripts/launch_list_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches_summary.txt 2>&1 || true
  tail -25 $OUT/${TAG}_launches_summary.txt
  gzip -f $OUT/${TAG}_launches.csv
fi
if has b1ab; then
  for prec in bf16 fp32_tc; do
    for env in "X=1" "DIM_SMALL_BATCH_GRAPH=1"; do
      echo "== vico_b1 $prec $env"
      env $env timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-leg --no-other-workloads --workload vico_b1 --precision $prec 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print(round(d['value']), 'frames/s', round(d['ms_per_step'],2), 'ms'); [print('   ',k) for k in d['kernels'][:4]]"
    done
  done
fi
if has gpab; then
  for e in "X=1" "DIM_TC_NO_GROUPED_PLANES=1"; do
    echo "== $e"
    env $e timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-workloads 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('bf16', round(d['value']), round(d['ms_per_step'],1), 'parity', round(d['fp32_parity_mode']['value']), round(d['fp32_parity_mode']['ms_per_step'],1)); [print('   ',k) for k in d['kernels'][:4]]"
  done
fi
if has lgen; then
  timeout 900 python -m pytest tests/test_listener_generator_gpu.py tests/test_slm_gpu.py tests/test_slmft_gpu.py tests/test_compat_gpu.py -x -q > $OUT/${TAG}_lgen.log 2>&1; echo "exit $?" >> $OUT/${TAG}_lgen.log
  tail -25 $OUT/${TAG}_lgen.log
fi
if has sanitize; then
  # memcheck over the round-2 kernels on small shapes (the 2^20-token argmin cases and the long tests are deselected: memcheck is ~30x slower)
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q \
      tests/test_decode_mk_gpu.py tests/test_loader_gpu.py tests/test_vqspeaker_gpu.py tests/test_slm_gpu.py tests/test_listener_generator_gpu.py \
      "tests/test_vq_argmin_tc_gpu.py" -k "not 1048576 and not 76800 and not small_batch_gemv" > $OUT/${TAG}_memcheck.log 2>&1
  echo "exit $?" >> $OUT/${TAG}_memcheck.log
  grep -E "ERROR SUMMARY|passed|failed|exit|Invalid|out of bounds" $OUT/${TAG}_memcheck.log | head -20
fi
if has q16ab; then
  for e in "X=1" "DIM_ATTN_QKV_FP32=1"; do
    echo "== $e"
    env $e timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-workloads --no-parity-leg 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print('bf16', round(d['value']), round(d['ms_per_step'],1)); [print('   ',k) for k in d['kernels'][:5]]"
  done
fi
if has mesh; then
  timeout 900 python -m pytest tests/test_speaker_mesh_gpu.py -q > $OUT/${TAG}_mesh.log 2>&1; echo "exit $?" >> $OUT/${TAG}_mesh.log
  tail -40 $OUT/${TAG}_mesh.log
  timeout 300 python scripts/lstm_timing.py > $OUT/${TAG}_lstm_timing.txt 2>&1; tail -12 $OUT/${TAG}_lstm_timing.txt
fi
if has pdlb1; then
  for env in "X=1" "DIM_PDL=1"; do
    echo "== vico_b1 bf16 $env"
    env $env timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-leg --no-other-workloads --workload vico_b1 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); print(round(d['value']), 'frames/s', round(d['ms_per_step'],2), 'ms'); [print('   ',k) for k in d['kernels'][:4]]"
  done
  DIM_PDL=1 timeout 900 python -m pytest tests/test_slmft_gpu.py -x -q > $OUT/${TAG}_slmft_pdl.log 2>&1; echo "exit $?" >> $OUT/${TAG}_slmft_pdl.log
  tail -5 $OUT/${TAG}_slmft_pdl.log
fi
if has sanitize2; then
  # memcheck + racecheck over the last kernels of the round: LSTM recurrence, alignment-free Linear, implicit-conv GEMM (small shapes)
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 --target-processes all python -m pytest -x -q \
      tests/test_speaker_mesh_gpu.py tests/test_conv_implicit_gpu.py tests/test_vqvae_gpu.py -k "not 256-12 and not 300-6 and not 8-138 and not full_size and not 70110" > $OUT/${TAG}_memcheck2.log 2>&1
  echo "exit $?" >> $OUT/${TAG}_memcheck2.log
  grep -E "ERROR SUMMARY|passed|failed|exit|Invalid|out of bounds" $OUT/${TAG}_memcheck2.log | sort | uniq -c | head -20
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest -x -q \
      tests/test_speaker_mesh_gpu.py -k "lstm and (1-27 or 33-50 or 5-40 or 3-1)" > $OUT/${TAG}_racecheck_lstm.log 2>&1
  echo "exit $?" >> $OUT/${TAG}_racecheck_lstm.log
  grep -E "RACECHECK SUMMARY|passed|failed|exit|hazard" $OUT/${TAG}_racecheck_lstm.log | sort | uniq -c | head -20
fi
if has ncuprefill; then
  # --set full captures of the prefill attention (bf16 operands, Dh = 64: the SLMFT encoders) and of one tcgen05 prefill GEMM
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:attn_prefill_mma --launch-skip 8 -c 2 -o $OUT/${TAG}_ncu_attn -f \
      python scripts/prefill_once.py > $OUT/${TAG}_ncu_attn.log 2>&1
  ncu -i $OUT/${TAG}_ncu_attn.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_attn.raw.csv 2>/dev/null
  ncu -i $OUT/${TAG}_ncu_attn.ncu-rep --page details > $OUT/${TAG}_ncu_attn.details.txt 2>/dev/null
  ncu -i $OUT/${TAG}_ncu_attn.ncu-rep --page source --csv --print-source sass > $OUT/${TAG}_ncu_attn.source.csv 2>/dev/null
  gzip -f $OUT/${TAG}_ncu_attn.source.csv
  ls -la $OUT/${TAG}_ncu_attn*; tail -3 $OUT/${TAG}_ncu_attn.log
fi

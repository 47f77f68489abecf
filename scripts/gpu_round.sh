#!/bin/bash
# One gpurun call: GPU parity tests, benches, the VQ-lookup roofline, and the ncu evidence (launch list + --set full captures).
# Usage (from the repo root on the GPU box):  bash scripts/gpu_round.sh [tag] [sections]
#   sections: any of  tests bench vq ncu_list ncu_full variants   (default: all but variants)
set -u
TAG=${1:-r01}
SECTIONS=${2:-"tests bench vq ncu_list ncu_full"}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $SECTIONS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1

if has tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1
  echo "tests exit $?" >> $OUT/${TAG}_tests.log
  tail -3 $OUT/${TAG}_tests.log
fi
if has bench; then
  timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err
  tail -c 600 $OUT/${TAG}_bench_bf16.json
  timeout 600 python bench.py --steps 3 --warmup 3 --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
  timeout 300 python bench.py --steps 3 --warmup 3 --precision fp32_tc --no-cpu-baseline > $OUT/${TAG}_bench_fp32tc.json 2> $OUT/${TAG}_bench_fp32tc.err
  timeout 300 python bench.py --steps 3 --warmup 3 --workload vico_b1 --precision fp32_tc --no-cpu-baseline > $OUT/${TAG}_bench_b1.json 2> $OUT/${TAG}_bench_b1.err
fi
if has variants; then
  DIM_NO_GROUPS=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-leg > $OUT/${TAG}_bench_bf16_nogroups.json 2>&1
  DIM_ATTN_NT=128 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-leg > $OUT/${TAG}_bench_bf16_nt128.json 2>&1
  DIM_ATTN_NT=64 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precision fp32_tc > $OUT/${TAG}_bench_fp32tc_nt64.json 2>&1
  DIM_PDL=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-leg > $OUT/${TAG}_bench_bf16_pdl.json 2>&1
fi
if has vq; then
  timeout 300 python scripts/vq_roofline.py > $OUT/${TAG}_vq_roofline.jsonl 2> $OUT/${TAG}_vq_roofline.err
  cat $OUT/${TAG}_vq_roofline.jsonl
fi
NCU="ncu --clock-control none"
BENCH_NCU="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-leg"
export DIM_BENCH_ALLOW_COLD=1 DIM_NO_GRAPH=1
if has ncu_list; then
  # every launch of the profiling-sized workload (same kernels, 64 clips x 48 frames), device time per launch
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $OUT/${TAG}_ncu_launches_mid_bf16.csv \
      $BENCH_NCU --workload mid > $OUT/${TAG}_ncu_list.log 2>&1
fi
if has ncu_full; then
  # --set full on the dominant kernels AT THE FULL WORKLOAD SIZE (B=256, T=300): decode attention deep into the decode loop
  # (self + cross), the skinny decode GEMMs, the big prefill GEMMs, layer norm; VQ kernels from the roofline script.
  timeout 900 $NCU --set full --import-source on -k regex:attn_decode -s 4200 -c 4 -o $OUT/${TAG}_ncu_attn_decode_bf16 -f \
      $BENCH_NCU > $OUT/${TAG}_ncu_full_attn.log 2>&1
  timeout 900 $NCU --set full -k regex:gemm_bf16_tcgen05 -s 9000 -c 8 -o $OUT/${TAG}_ncu_gemm_decode_bf16 -f \
      $BENCH_NCU > $OUT/${TAG}_ncu_full_gemm.log 2>&1
  timeout 900 $NCU --set full -k regex:gemm_bf16_tcgen05 -s 1 -c 8 -o $OUT/${TAG}_ncu_gemm_prefill -f \
      $BENCH_NCU > $OUT/${TAG}_ncu_full_gemm2.log 2>&1
  timeout 900 $NCU --set full -k regex:"attn_prefill|layer_norm|instance_norm" -s 20 -c 6 -o $OUT/${TAG}_ncu_prefill_misc -f \
      $BENCH_NCU > $OUT/${TAG}_ncu_full_misc.log 2>&1
  VQ_NCU=1 timeout 600 $NCU --set full --import-source on -k regex:"vq_gather|vq_argmin" -c 6 -o $OUT/${TAG}_ncu_vq -f \
      python scripts/vq_roofline.py > $OUT/${TAG}_ncu_full_vq.log 2>&1
fi
ls -la $OUT | tail -30

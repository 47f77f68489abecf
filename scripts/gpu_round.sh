#!/bin/bash
# One gpurun call: GPU parity tests, benches, the VQ-lookup roofline, and the ncu evidence (launch list + --set full captures).
# Usage (from the repo root on the GPU box):  bash scripts/gpu_round.sh [tag] [sections]
#   sections: any of  tests bench variants vq ncu_list ncu_full   (default: all but variants); vq also runs the decode-attention
#   roofline and the in-graph per-kernel cost script
# gpurun copies gpurun_out/ back only when it is <= 64 MiB: every .ncu-rep is reduced to its raw-page CSV on the box and
# dropped when large; the directory is pruned to < 48 MiB at the end.
set -u
TAG=${1:-r01}
SECTIONS=${2:-"tests bench vq ncu_list ncu_full"}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $SECTIONS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
brief() { python - "$1" <<'EOF'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    r = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], "value=%.0f ms=%.1f e2e=%.0f" % (d["value"], d["ms_per_step"], d.get("e2e", {}).get("value", 0)),
          "parity=%.0f" % d.get("fp32_parity_mode", {}).get("value", 0), "launches=%s" % d.get("gpu_launches"),
          "roof=%s %.3f" % (r.get("kernel"), r.get("frac", 0)), "clocks=%s" % d.get("clocks"))
    for k in d.get("kernels", [])[:7]:
        print("    ", k)
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
EOF
}
bench() { # name, [VAR=val ...] cmd args
  local name=$1; shift
  timeout 600 env "$@" > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_bench_${name}.err
  brief $OUT/${TAG}_bench_${name}.json
}

if has tests; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1
  echo "tests exit $?" >> $OUT/${TAG}_tests.log
  tail -15 $OUT/${TAG}_tests.log
fi
if has bench; then
  bench bf16 python bench.py --steps 3 --warmup 3
  timeout 600 python bench.py --steps 3 --warmup 3 --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
  cut -c1-400 $OUT/${TAG}_bench_reference.json
  bench fp32tc python bench.py --steps 3 --warmup 3 --precision fp32_tc --no-cpu-baseline
  bench b1 python bench.py --steps 3 --warmup 3 --workload vico_b1 --precision fp32_tc --no-cpu-baseline
fi
if has variants; then
  V="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-leg"
  bench bf16_nogroups DIM_NO_GROUPS=1 $V
  bench bf16_g64x4 DIM_GROUP_ROWS=64 DIM_MAX_GROUPS=4 $V
  bench bf16_nt128 DIM_ATTN_NT=128 $V
  bench bf16_ffma_prefill DIM_ATTN_PREFILL=ffma $V
  bench bf16_pdl DIM_PDL=1 $V
  bench bf16_pdl_nogroups DIM_PDL=1 DIM_NO_GROUPS=1 $V
fi
if has vq; then
  timeout 300 python scripts/vq_roofline.py > $OUT/${TAG}_vq_roofline.jsonl 2> $OUT/${TAG}_vq_roofline.err
  cut -c1-330 $OUT/${TAG}_vq_roofline.jsonl
  timeout 300 python scripts/attn_roofline.py > $OUT/${TAG}_attn_roofline.jsonl 2> $OUT/${TAG}_attn_roofline.err
  grep default $OUT/${TAG}_attn_roofline.jsonl | cut -c1-260
  timeout 300 python scripts/graph_gap.py 256 > $OUT/${TAG}_graph_gap.txt 2>&1
  cat $OUT/${TAG}_graph_gap.txt
fi
NCU="ncu --clock-control none"
BENCH_NCU="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-leg"
export DIM_BENCH_ALLOW_COLD=1 DIM_NO_GRAPH=1
if has ncu_list; then
  # every launch of the profiling-sized workload (same kernels, 64 clips x 48 frames), device time per launch
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $OUT/${TAG}_ncu_launches_mid_bf16.csv \
      $BENCH_NCU --workload mid > $OUT/${TAG}_ncu_list.log 2>&1
  gzip -f $OUT/${TAG}_ncu_launches_mid_bf16.csv
fi
full() { # name, kernel regex, skip, count, cmd...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 900 $NCU --set full -k regex:"$rx" -s $skip -c $cnt -o $OUT/${TAG}_ncu_$name -f "$@" > $OUT/${TAG}_ncu_full_$name.log 2>&1
  ncu -i $OUT/${TAG}_ncu_$name.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_$name.raw.csv 2>/dev/null
  local sz=$(stat -c %s $OUT/${TAG}_ncu_$name.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 9000000 ]; then rm -f $OUT/${TAG}_ncu_$name.ncu-rep; fi
  echo "ncu $name: rep $sz bytes, csv $(wc -l < $OUT/${TAG}_ncu_$name.raw.csv) lines"
}
if has ncu_full; then
  # --set full on the dominant kernels AT THE FULL WORKLOAD SIZE (B=256, T=300): decode attention deep into the decode loop
  # (self + cross), the skinny decode GEMMs, the big prefill GEMMs, prefill attention / norms; VQ kernels from the roofline script.
  full attn_decode_bf16 attn_decode 4200 4 $BENCH_NCU
  full gemm_decode_bf16 gemm_bf16_tcgen05 9000 6 $BENCH_NCU
  full gemm_prefill gemm_bf16_tcgen05 1 6 $BENCH_NCU
  full prefill_misc "attn_prefill|layer_norm|instance_norm" 20 6 $BENCH_NCU
  VQ_NCU=1 full vq "vq_gather|vq_argmin" 0 9 python scripts/vq_roofline.py
fi
# keep the directory under the copy-back limit
while [ "$(du -sm $OUT | cut -f1)" -gt 48 ]; do
  big=$(ls -S $OUT | head -1); echo "pruning $big"; rm -f "$OUT/$big"
done
du -sh $OUT

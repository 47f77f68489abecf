#!/bin/bash
# A/B of the persistent prefill GEMM (DIM_GEMM_PERSIST=0 keeps the one-tile-per-CTA kernel): bash scripts/r02_ab_persist.sh TAG
TAG=${1:-r02ab}
OUT=gpurun_out
mkdir -p $OUT
for m in 1 0; do
  DIM_GEMM_PERSIST=$m timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_persist$m.json 2> $OUT/${TAG}_bench_persist$m.err
  python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench_persist$m.json"))
print("persist=$m value", round(d["value"]), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "parity", round(d["fp32_parity_mode"]["value"]), round(d["fp32_parity_mode"]["ms_per_step"],1))
for k in d["kernels"][:4]: print("   ", k["kernel"], k["launches"], k["ms"], k.get("TFLOP/s"))
PY
done

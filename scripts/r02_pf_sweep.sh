#!/bin/bash
# sweep of the K/V L2 prefetch budget (DIM_MK_PF_MB) and the weight L2 policy of the persistent decode kernel
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_decode_mk_gpu.py -x -q 2>&1 | tail -2
for cfg in "0 0.7" "48 0.7" "96 0.7" "96 0" "144 0" "200 0"; do
  set -- $cfg
  echo "== DIM_MK_PF_MB=$1 DIM_L2_WEIGHT_KEEP=$2"
  for prec in ${PRECS:-bf16}; do
    DIM_MK_PF_MB=$1 DIM_L2_WEIGHT_KEEP=$2 timeout 300 python scripts/decode_trace.py 256 300 $prec 2>&1 | grep -v "  phase 1\|  phase 2\|  phase 3\|  phase 4"
  done
done

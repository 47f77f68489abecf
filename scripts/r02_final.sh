#!/bin/bash
# Final evidence of a round-2 session in ONE gpurun call: GPU tests, smoke(), the default bench line, the reference arm, and the ncu
# launch list of the same bench command.  bash scripts/r02_final.sh TAG
set -u
TAG=${1:-r02fin}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "exit $?" >> $OUT/${TAG}_tests.log
tail -4 $OUT/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "exit $?" >> $OUT/${TAG}_smoke.log
tail -3 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err; echo "bench exit $?"
cut -c1-700 $OUT/${TAG}_bench_bf16.json
timeout 900 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference exit $?"
cut -c1-500 $OUT/${TAG}_bench_reference.json
DIM_BENCH_ALLOW_COLD=1 timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum -c 3000 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-leg > $OUT/${TAG}_ncu_list.log 2>&1
python scripts/launch_list_summary.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launch_list_summary.txt 2>&1
head -14 $OUT/${TAG}_launch_list_summary.txt
gzip -f $OUT/${TAG}_launches.csv

#!/bin/bash
# ncu passes of bench.py on one B200 (never a multi-rank command).  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh TAG'
TAG=${1:-prof}
mkdir -p gpurun_out
# every launch with its device time / DRAM bytes / instructions (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --frames 32 --no-cpu-baseline --no-e2e --no-parity --no-configs --no-sustained > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu launches exit $?"
# the two transform kernels, full set, source-level
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"encode_fast|decode_fast" -s 4 -c 2 \
    -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --frames 32 --no-cpu-baseline --no-e2e --no-parity --no-configs --no-sustained > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"

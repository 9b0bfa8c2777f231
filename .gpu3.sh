set -x
timeout 600 python bench.py > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err; echo "bench exit $?"; cat gpurun_out/r1c_bench.json; tail -5 gpurun_out/r1c_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1c_ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"encode|decode" -s 4 -c 2 -o gpurun_out/r1c_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1c_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 300 python bench.py --frames 1 --no-cpu-baseline --no-e2e
timeout 300 python bench.py --frames 32 --no-cpu-baseline --no-e2e

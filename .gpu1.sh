set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc; lscpu | grep -E "Model name|^CPU\(s\)"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r1_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/r1_smoke.log
timeout 600 python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench exit $?"; cat gpurun_out/r1_bench.json; tail -5 gpurun_out/r1_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_ncu_bench.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"encode_kernel|decode_kernel" -s 4 -c 4 -o gpurun_out/r1_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out

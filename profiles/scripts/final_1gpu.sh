# End-of-round pass on one B200: tests, ncu capture of the run kernel (bench.py quotes its counters), launch list of the
# bench command, the bench line, run kernel vs per-step launches, back-to-back determinism.
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; tail -2 gpurun_out/r2_final_pytest.log
bash profiles/run_profile_run_kernel.sh r2
cp gpurun_out/r2_ncu_c1_run_summary.csv profiles/r2_ncu_c1_run_summary.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-stress --no-extra > gpurun_out/bench_under_ncu_r2.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 600 gpurun_out/r2_bench_1gpu.json
timeout 300 python profiles/run_vs_step.py 200 1024 2048 4096 > gpurun_out/r2_run_vs_step.log 2>&1; cat gpurun_out/r2_run_vs_step.log
timeout 300 python profiles/stress_run_kernel.py 300 > gpurun_out/r2_stress_run_kernel.log 2>&1; tail -3 gpurun_out/r2_stress_run_kernel.log

# Single-GPU end-of-round pass after the shared records (order >= 5, in kernel instantiations of their own) went in:
# tests, the bench line, the step kernel with the machine full.
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; tail -2 gpurun_out/r2_final_pytest.log
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; tail -c 200 gpurun_out/r2_bench_1gpu.json
timeout 200 python profiles/throughput.py 6 16384 > gpurun_out/r2_throughput_o6.log 2>&1; tail -1 gpurun_out/r2_throughput_o6.log
timeout 200 python profiles/throughput.py 4 1024 16384 131072 1048576 > gpurun_out/r2_throughput_o4.log 2>&1; tail -2 gpurun_out/r2_throughput_o4.log

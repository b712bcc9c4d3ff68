set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_pytest.log 2>&1; tail -3 gpurun_out/t_pytest.log
timeout 300 python profiles/run_vs_step.py 200 1024 > gpurun_out/t_run_vs_step_ll.log 2>&1; cat gpurun_out/t_run_vs_step_ll.log
QIW_RUN_BARRIER=1 timeout 300 python profiles/run_vs_step.py 200 1024 > gpurun_out/t_run_vs_step_barrier.log 2>&1; cat gpurun_out/t_run_vs_step_barrier.log
timeout 300 python profiles/throughput_block.py 3 1024 4096 > gpurun_out/t_throughput_block.log 2>&1; tail -3 gpurun_out/t_throughput_block.log
timeout 300 python profiles/throughput.py 4 1024 131072 > gpurun_out/t_throughput_o4.log 2>&1; tail -3 gpurun_out/t_throughput_o4.log
timeout 300 python profiles/stress_run_kernel.py 100 > gpurun_out/t_stress.log 2>&1; tail -3 gpurun_out/t_stress.log
timeout 600 python bench.py > gpurun_out/t_bench_1gpu.json 2> gpurun_out/t_bench_1gpu.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/t_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['pass'], d['roofline']['frac'], d['e2e'].get('host_stepped'), d['other_configs']['c4_two_band']['device_ms'])
PY

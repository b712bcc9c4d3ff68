# End-of-round pass on one 8-GPU box: the bench line at 8, 4 and 2 GPUs (C1 weak scaling, C5 strong scaling, parity
# against the oracle inside) and the multi-GPU parity check on all 8.
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29520 + $1)) \
          bench.py --gpus $1 --steps 5 --warmup 3 > gpurun_out/r2_bench_$1gpu_peer.json 2> gpurun_out/r2_bench_$1gpu.err; }
run 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tests/multigpu_check.py > gpurun_out/r2_multigpu_check_8gpu.log 2>&1
tail -1 gpurun_out/r2_multigpu_check_8gpu.log
run 4
run 2
python - <<'PY'
import json
for n in (2, 4, 8):
    try:
        d = json.loads(open('gpurun_out/r2_bench_%dgpu_peer.json' % n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'], 3), '%.3e' % d['value'], '%.3e' % d['e2e']['value'], d['parity']['pass'],
              round(d['stress_c5_step']['step_ms_device_max'], 1), d['stress_c5_step']['parity_pass'], round(d['roofline']['frac'], 3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(n, 'failed', e)
PY

# A/B of library builds on N GPUs: the README run, weak scaling (bench.py without the side sections), then the
# multi-GPU parity check of the library in the tree.  usage: bash profiles/scripts/multi_ab.sh N v1 v5 ...
N=$1; shift
for v in "$@"; do
  for rep in 1 2; do
    QIW_LIB=$PWD/variants/$v.so timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 5 --warmup 3 --no-stress --no-extra --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['n_gpus'], 'ms_per_step', round(d['ms_per_step'],3), 'e2e ms', round(d['diagram_evals_per_step']/d['e2e']['value']*1e3,3))"
  done
done
[ -n "$CHECK" ] && timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/multigpu_check.py 2>&1 | tail -2
[ -n "$CHECK" ] && QIW_FORCE_COMPLEX=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/multigpu_check.py 2>&1 | tail -2

# round 2, 2-GPU call: multi-GPU tests (both fused exchanges) + bench with the peer-to-peer and the in-switch exchange
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_multigpu.py -x -q -m gpu < /dev/null) > gpurun_out/tests_n2.log 2>&1; tail -3 gpurun_out/tests_n2.log
for ex in peer multimem; do
  R3D_SHARDED_ADAM_EXCHANGE=$ex timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 < /dev/null > gpurun_out/r02_bench_n2_$ex.json 2> gpurun_out/r02_bench_n2_$ex.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2_$ex.json')); print('$ex', d['value'], d['ms_per_step'], d['e2e']['value'], d['phases_ms'], d['parity_check'], d['config']['exchange'])"
  tail -2 gpurun_out/r02_bench_n2_$ex.err
done

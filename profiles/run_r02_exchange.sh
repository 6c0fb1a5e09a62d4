# round 2: launch-shape sweep of the fused in-switch exchange + optimizer on 8 GPUs (c3-sized grid, 1.95 GB)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for u in 4 8 2; do
  R3D_MULTIMEM_UNROLL=$u timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2956$u profiles/exchange_bench.py 2>/dev/null | grep '^{' >> gpurun_out/r02_exchange_sweep.jsonl
done
cat gpurun_out/r02_exchange_sweep.jsonl

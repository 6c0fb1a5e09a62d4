# round 2, call G (2 GPUs): fused reduce-scatter/Adam/all-gather parity test, 1- and 2-GPU bench lines with the optimizer in the step
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q) > gpurun_out/tests_mg.log 2>&1; tail -5 gpurun_out/tests_mg.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 1500 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_fused.json 2> gpurun_out/r02_bench_n2_fused.err; tail -c 1200 gpurun_out/r02_bench_n2_fused.json; tail -3 gpurun_out/r02_bench_n2_fused.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl > gpurun_out/r02_bench_n2_nccl.json 2> gpurun_out/r02_bench_n2_nccl.err; tail -c 600 gpurun_out/r02_bench_n2_nccl.json; tail -3 gpurun_out/r02_bench_n2_nccl.err

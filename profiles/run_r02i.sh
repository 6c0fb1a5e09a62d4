# round 2, call I (8 GPUs): 2-GPU parity tests, bench at N = 8 / 4 (fused in-switch optimizer), N = 8 with NCCL, c5 at N = 8
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q) > gpurun_out/tests_mg.log 2>&1; tail -3 gpurun_out/tests_mg.log
run() { n=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
run 8 29521 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8_fused.json 2> gpurun_out/r02_bench_n8_fused.err; tail -c 400 gpurun_out/r02_bench_n8_fused.json; tail -2 gpurun_out/r02_bench_n8_fused.err
run 4 29522 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4_fused.json 2> gpurun_out/r02_bench_n4_fused.err; tail -c 400 gpurun_out/r02_bench_n4_fused.json; tail -2 gpurun_out/r02_bench_n4_fused.err
run 8 29523 --steps 10 --warmup 3 --exchange nccl > gpurun_out/r02_bench_n8_nccl.json 2> gpurun_out/r02_bench_n8_nccl.err; tail -c 300 gpurun_out/r02_bench_n8_nccl.json; tail -2 gpurun_out/r02_bench_n8_nccl.err
run 8 29524 --steps 10 --warmup 3 --exchange nvls > gpurun_out/r02_bench_n8_nvls_allreduce.json 2> gpurun_out/r02_bench_n8_nvls_allreduce.err; tail -c 300 gpurun_out/r02_bench_n8_nvls_allreduce.json; tail -2 gpurun_out/r02_bench_n8_nvls_allreduce.err
run 8 29525 --steps 5 --warmup 3 --workload c5_512cube_deg3_1600px_512spp > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err; tail -c 600 gpurun_out/r02_bench_c5_n8.json; tail -3 gpurun_out/r02_bench_c5_n8.err

# round 2, call H (2 GPUs): full GPU suite on the product build and on the A/B build, 2-GPU tests, traffic measurement, bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_prod.log 2>&1; tail -4 gpurun_out/tests_prod.log
(time R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/libr3d_b200_ab.so timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "cooperative_and_per_ray or contribution or mask") > gpurun_out/tests_ab.log 2>&1; tail -4 gpurun_out/tests_ab.log
timeout 600 python profiles/measure_traffic.py > gpurun_out/traffic.log 2>&1; tail -2 gpurun_out/traffic.log; cp profiles/traffic.json gpurun_out/traffic.json
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 700 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_fused.json 2> gpurun_out/r02_bench_n2_fused.err; tail -c 900 gpurun_out/r02_bench_n2_fused.json; tail -3 gpurun_out/r02_bench_n2_fused.err

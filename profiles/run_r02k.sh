set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q) > gpurun_out/tests_mg.log 2>&1; grep -n "^E  " gpurun_out/tests_mg.log | tail -5; tail -2 gpurun_out/tests_mg.log
run() { n=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
run 2 29541 --steps 10 --warmup 3 --strong > gpurun_out/r02_bench_n2_strong.json 2> gpurun_out/r02_bench_n2_strong.err; tail -c 300 gpurun_out/r02_bench_n2_strong.json; tail -2 gpurun_out/r02_bench_n2_strong.err
run 2 29542 --steps 10 --warmup 3 --exchange nvls > gpurun_out/r02_bench_n2_nvls.json 2> gpurun_out/r02_bench_n2_nvls.err; tail -c 300 gpurun_out/r02_bench_n2_nvls.json; tail -2 gpurun_out/r02_bench_n2_nvls.err
run 2 29543 --steps 10 --warmup 3 --no-optimizer > gpurun_out/r02_bench_n2_noopt.json 2> gpurun_out/r02_bench_n2_noopt.err; tail -c 300 gpurun_out/r02_bench_n2_noopt.json; tail -2 gpurun_out/r02_bench_n2_noopt.err

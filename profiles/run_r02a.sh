# round 2, call A: FFMA2 micro-benchmark, GPU tests, A/B of the lane-group (0) vs warp-specialised (32) forward, ncu of the latter
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 60 profiles/ubench/ffma2.bin > gpurun_out/r02_ubench_ffma2.json 2>&1; cat gpurun_out/r02_ubench_ffma2.json
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/tests.log 2>&1; tail -6 gpurun_out/tests.log
timeout 300 python profiles/ab_kernels.py --variants 0,32,0,32 --iters 10 > gpurun_out/ab.json 2> gpurun_out/ab.err; tail -6 gpurun_out/ab.err
timeout 300 python profiles/ab_kernels.py --variants 0,32 --iters 10 --density-shift 0.9 > gpurun_out/ab_sparse.json 2> gpurun_out/ab_sparse.err; tail -3 gpurun_out/ab_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_ws -s 3 -c 1 -f -o gpurun_out/r02_ws1 python profiles/ab_kernels.py --variants 32 --iters 1 > gpurun_out/ncu_ws1.log 2>&1; tail -3 gpurun_out/ncu_ws1.log

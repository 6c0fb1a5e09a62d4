# round 2, call C: producer-side prefetch (L2 / L1) for the ws forward, first run of the ws backward
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_features.py tests/test_gpu_parity.py -m gpu -x -q) > gpurun_out/tests.log 2>&1; tail -6 gpurun_out/tests.log
timeout 300 python profiles/ab_kernels.py --variants 0,96,352,608,128,224 --iters 10 > gpurun_out/ab_c.json 2> gpurun_out/ab_c.err; tail -7 gpurun_out/ab_c.err
timeout 300 python profiles/ab_kernels.py --variants 0,128,224 --iters 10 --density-shift 0.9 > gpurun_out/ab_c_sparse.json 2> gpurun_out/ab_c_sparse.err; tail -4 gpurun_out/ab_c_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd_ws -s 3 -c 1 -f -o gpurun_out/r02_wsb1 python profiles/ab_kernels.py --variants 128 --iters 1 > gpurun_out/ncu_wsb1.log 2>&1; tail -2 gpurun_out/ncu_wsb1.log

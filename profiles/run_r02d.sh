# round 2, call D: ws forward with 4 stages / lag 3, dual-stream consumer (2 CTAs/SM, 128 regs), with / without L2 prefetch
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "cooperative_and_per_ray or config3 or jitter") > gpurun_out/tests.log 2>&1; tail -4 gpurun_out/tests.log
timeout 300 python profiles/ab_kernels.py --variants 0,96,352,1120,1376,1376 --iters 10 > gpurun_out/ab_d.json 2> gpurun_out/ab_d.err; tail -7 gpurun_out/ab_d.err
timeout 300 python profiles/ab_kernels.py --variants 0,352,1376 --iters 10 --density-shift 0.9 > gpurun_out/ab_d_sparse.json 2> gpurun_out/ab_d_sparse.err; tail -4 gpurun_out/ab_d_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_ws -s 3 -c 1 -f -o gpurun_out/r02_ws3_dq python profiles/ab_kernels.py --variants 1376 --iters 1 > gpurun_out/ncu_ws3.log 2>&1; tail -2 gpurun_out/ncu_ws3.log

# round 2, call B: ws forward with density quads / run-merged gathers / cell-sorted publish: tests, A/B, ncu
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/tests.log 2>&1; tail -6 gpurun_out/tests.log
timeout 300 python profiles/ab_kernels.py --variants 0,32,96,32,96 --iters 10 > gpurun_out/ab_q1.json 2> gpurun_out/ab_q1.err; tail -6 gpurun_out/ab_q1.err
R3D_DENSITY_QUADS=0 timeout 300 python profiles/ab_kernels.py --variants 32,96 --iters 10 > gpurun_out/ab_q0.json 2> gpurun_out/ab_q0.err; tail -3 gpurun_out/ab_q0.err
timeout 300 python profiles/ab_kernels.py --variants 0,32,96 --iters 10 --density-shift 0.9 > gpurun_out/ab_sparse.json 2> gpurun_out/ab_sparse.err; tail -4 gpurun_out/ab_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_ws -s 3 -c 1 -f -o gpurun_out/r02_ws2_v32 python profiles/ab_kernels.py --variants 32 --iters 1 > gpurun_out/ncu_ws2a.log 2>&1; tail -2 gpurun_out/ncu_ws2a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_ws -s 3 -c 1 -f -o gpurun_out/r02_ws2_v96 python profiles/ab_kernels.py --variants 96 --iters 1 > gpurun_out/ncu_ws2b.log 2>&1; tail -2 gpurun_out/ncu_ws2b.log

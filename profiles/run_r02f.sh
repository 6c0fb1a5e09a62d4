# round 2, call F: lane-group forward with cell-sorted publish (hardware coalescing of same-cell gathers)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "cooperative_and_per_ray") > gpurun_out/tests.log 2>&1; tail -4 gpurun_out/tests.log
timeout 300 python profiles/ab_kernels.py --variants 0,2048,0,2048 --iters 10 > gpurun_out/ab_f.json 2> gpurun_out/ab_f.err; tail -5 gpurun_out/ab_f.err
timeout 300 python profiles/ab_kernels.py --variants 0,2048 --iters 10 --density-shift 0.9 > gpurun_out/ab_f_sparse.json 2> gpurun_out/ab_f_sparse.err; tail -3 gpurun_out/ab_f_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_group -s 3 -c 1 -f -o gpurun_out/r02_group_sort python profiles/ab_kernels.py --variants 2048 --iters 1 > gpurun_out/ncu_gs.log 2>&1; tail -2 gpurun_out/ncu_gs.log

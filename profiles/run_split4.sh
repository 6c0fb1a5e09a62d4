# usage (GPU box): bash profiles/run_split4.sh -- agreement test on the measurement build + ncu captures of the two-kernel forward
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
export R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/libr3d_b200_ab.so
(timeout 600 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "agree" < /dev/null) > gpurun_out/split_tests.log 2>&1
tail -3 gpurun_out/split_tests.log
R3D_SPLIT_MODE=9 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_gather -s 3 -c 1 -f -o gpurun_out/r02_split_gather_nomath python profiles/ab_kernels.py --variants 32768 --iters 1 < /dev/null > gpurun_out/ncu_nomath.log 2>&1
tail -2 gpurun_out/ncu_nomath.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_probe -s 3 -c 1 -f -o gpurun_out/r02_split_probe python profiles/ab_kernels.py --variants 32768 --iters 1 < /dev/null > gpurun_out/ncu_probe.log 2>&1
tail -2 gpurun_out/ncu_probe.log

# usage (on the GPU box): bash profiles/run_split.sh  -- A/B of the two-kernel forward (measurement build)
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
export R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/libr3d_b200_ab.so
(timeout 600 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "agree" < /dev/null) > gpurun_out/split_tests.log 2>&1
tail -5 gpurun_out/split_tests.log
timeout 300 python profiles/ab_kernels.py --variants ${VARIANTS:-0,32768} --iters 10 < /dev/null > gpurun_out/split_ab.json 2> gpurun_out/split_ab.err
tail -4 gpurun_out/split_ab.err
timeout 300 python profiles/ab_kernels.py --variants ${VARIANTS:-0,32768} --iters 10 --density-shift 0.9 < /dev/null > gpurun_out/split_ab_sparse.json 2> gpurun_out/split_ab_sparse.err
tail -4 gpurun_out/split_ab_sparse.err
if [ -n "$NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:render_ -c 40 --csv --log-file gpurun_out/split_launches.csv python profiles/ab_kernels.py --variants 32768 --iters 2 --warmup 1 < /dev/null > gpurun_out/split_ncu.log 2>&1
  grep -c render_ gpurun_out/split_launches.csv
fi

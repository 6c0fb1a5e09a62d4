# round 2, call E: FFMA2 in the cooperative backward's member sweep; ws forward at 4 CTAs/SM (64 regs) vs 3
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libr3d_b200.so libr3d_b200_ws4.so libr3d_b200_ws3s3.so; do
  R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/$lib timeout 300 python profiles/ab_kernels.py --variants 0,352 --iters 10 > gpurun_out/ab_e_$lib.json 2> gpurun_out/ab_e_$lib.err
  tail -2 gpurun_out/ab_e_$lib.err
done
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training_parity.py tests/test_gpu_dual_render.py -m gpu -x -q) > gpurun_out/tests.log 2>&1; tail -4 gpurun_out/tests.log

# usage (on the GPU box): ARGS="..." bash profiles/run_split2.sh  -- one ab_kernels run on the measurement build
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
export R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/libr3d_b200_ab.so
timeout 400 python profiles/ab_kernels.py $ARGS < /dev/null > gpurun_out/${OUT:-split2}.json 2> gpurun_out/${OUT:-split2}.err
tail -8 gpurun_out/${OUT:-split2}.err

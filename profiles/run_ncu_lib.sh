# usage (GPU box): bash profiles/run_ncu_lib.sh <lib.so> <kernel-regex> <out-name>
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/$1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/$3 python profiles/ab_kernels.py --variants 0 --iters 1 > gpurun_out/ncu_$3.log 2>&1
tail -3 gpurun_out/ncu_$3.log

# usage (GPU box): bash profiles/run_ncu_variant.sh <variant> <kernel-regex> <out-name>   (measurement build)
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/libr3d_b200_ab.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/$3 python profiles/ab_kernels.py --variants $1 --iters 1 < /dev/null > gpurun_out/ncu_$3.log 2>&1
tail -3 gpurun_out/ncu_$3.log

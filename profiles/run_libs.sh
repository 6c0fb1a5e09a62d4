# usage (GPU box): bash profiles/run_libs.sh "<ab args>" lib1.so lib2.so ...   -- A/B of differently compiled libraries
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
args="$1"; shift
for lib in "$@"; do
  R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/$lib timeout 300 python profiles/ab_kernels.py $args > gpurun_out/ab_$lib.json 2> gpurun_out/ab_$lib.err
  tail -4 gpurun_out/ab_$lib.err
done

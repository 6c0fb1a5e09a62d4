# usage (on the GPU box): bash profiles/run_ab.sh "<pytest args or empty>" "<ab_kernels args>" [ncu-kernel-regex out-name]
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
if [ -n "$1" ]; then
  (time timeout 900 python -m pytest $1 -m gpu -x -q) > gpurun_out/tests.log 2>&1
  tail -8 gpurun_out/tests.log
fi
if [ -n "$2" ]; then
  timeout 600 python profiles/ab_kernels.py $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  tail -12 gpurun_out/ab.err
fi
if [ -n "$3" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -f -o gpurun_out/$4 python profiles/ab_kernels.py --variants 0 --iters 1 > gpurun_out/ncu_$4.log 2>&1
  tail -3 gpurun_out/ncu_$4.log
fi

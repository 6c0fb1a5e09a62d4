# usage (GPU box): bash profiles/run_env.sh "<ab args>" "ENV=val" "ENV=val2" ...   -- A/B over environment settings
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
args="$1"; shift
for e in "$@"; do
  env $e timeout 300 python profiles/ab_kernels.py $args > "gpurun_out/ab_$e.json" 2> "gpurun_out/ab_$e.err"
  tail -4 "gpurun_out/ab_$e.err"
done

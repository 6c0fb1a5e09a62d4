# usage (GPU box): bash profiles/run_ncu_adam.sh -- ncu --set full of the fused dense Adam kernel inside the bench step (4th launch = the feature tensor of the second step)
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:adam_kernel -s 3 -c 1 -f -o gpurun_out/r02_final_adam python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-baseline < /dev/null > gpurun_out/ncu_final_adam.log 2>&1
tail -2 gpurun_out/ncu_final_adam.log

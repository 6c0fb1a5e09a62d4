# usage (GPU box): bash profiles/run_split5.sh -- gather micro-benchmark + the gather kernel without arithmetic (full-width loads)
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 120 profiles/ubench/gather.bin > gpurun_out/r02_ubench_gather.json 2>&1; cat gpurun_out/r02_ubench_gather.json
export R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/libr3d_b200_ab.so
R3D_SPLIT_MODE=9 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:render_ -c 12 --csv --log-file gpurun_out/r02_split_nomath_launches.csv python profiles/ab_kernels.py --variants 32768 --iters 2 --warmup 1 < /dev/null > gpurun_out/split_ncu.log 2>&1
grep render_ gpurun_out/r02_split_nomath_launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
R3D_SPLIT_MODE=9 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_gather -s 3 -c 1 -f -o gpurun_out/r02_split_gather_nomath python profiles/ab_kernels.py --variants 32768 --iters 1 < /dev/null > gpurun_out/ncu_nomath.log 2>&1
tail -1 gpurun_out/ncu_nomath.log

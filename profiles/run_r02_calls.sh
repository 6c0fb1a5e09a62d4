# The gpurun command scripts of round 2, in the order they were run (each was `gpurun -- 'bash profiles/run_r02X.sh'`).
# Kept as the record of how the profiles/r02_* files were produced; run one section at a time.

### run_r02a.sh
# round 2, call A: FFMA2 micro-benchmark, GPU tests, A/B of the lane-group (0) vs warp-specialised (32) forward, ncu of the latter
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 60 profiles/ubench/ffma2.bin > gpurun_out/r02_ubench_ffma2.json 2>&1; cat gpurun_out/r02_ubench_ffma2.json
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/tests.log 2>&1; tail -6 gpurun_out/tests.log
timeout 300 python profiles/ab_kernels.py --variants 0,32,0,32 --iters 10 > gpurun_out/ab.json 2> gpurun_out/ab.err; tail -6 gpurun_out/ab.err
timeout 300 python profiles/ab_kernels.py --variants 0,32 --iters 10 --density-shift 0.9 > gpurun_out/ab_sparse.json 2> gpurun_out/ab_sparse.err; tail -3 gpurun_out/ab_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_ws -s 3 -c 1 -f -o gpurun_out/r02_ws1 python profiles/ab_kernels.py --variants 32 --iters 1 > gpurun_out/ncu_ws1.log 2>&1; tail -3 gpurun_out/ncu_ws1.log

### run_r02b.sh
# round 2, call B: ws forward with density quads / run-merged gathers / cell-sorted publish: tests, A/B, ncu
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/tests.log 2>&1; tail -6 gpurun_out/tests.log
timeout 300 python profiles/ab_kernels.py --variants 0,32,96,32,96 --iters 10 > gpurun_out/ab_q1.json 2> gpurun_out/ab_q1.err; tail -6 gpurun_out/ab_q1.err
R3D_DENSITY_QUADS=0 timeout 300 python profiles/ab_kernels.py --variants 32,96 --iters 10 > gpurun_out/ab_q0.json 2> gpurun_out/ab_q0.err; tail -3 gpurun_out/ab_q0.err
timeout 300 python profiles/ab_kernels.py --variants 0,32,96 --iters 10 --density-shift 0.9 > gpurun_out/ab_sparse.json 2> gpurun_out/ab_sparse.err; tail -4 gpurun_out/ab_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_ws -s 3 -c 1 -f -o gpurun_out/r02_ws2_v32 python profiles/ab_kernels.py --variants 32 --iters 1 > gpurun_out/ncu_ws2a.log 2>&1; tail -2 gpurun_out/ncu_ws2a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_ws -s 3 -c 1 -f -o gpurun_out/r02_ws2_v96 python profiles/ab_kernels.py --variants 96 --iters 1 > gpurun_out/ncu_ws2b.log 2>&1; tail -2 gpurun_out/ncu_ws2b.log

### run_r02c.sh
# round 2, call C: producer-side prefetch (L2 / L1) for the ws forward, first run of the ws backward
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_features.py tests/test_gpu_parity.py -m gpu -x -q) > gpurun_out/tests.log 2>&1; tail -6 gpurun_out/tests.log
timeout 300 python profiles/ab_kernels.py --variants 0,96,352,608,128,224 --iters 10 > gpurun_out/ab_c.json 2> gpurun_out/ab_c.err; tail -7 gpurun_out/ab_c.err
timeout 300 python profiles/ab_kernels.py --variants 0,128,224 --iters 10 --density-shift 0.9 > gpurun_out/ab_c_sparse.json 2> gpurun_out/ab_c_sparse.err; tail -4 gpurun_out/ab_c_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd_ws -s 3 -c 1 -f -o gpurun_out/r02_wsb1 python profiles/ab_kernels.py --variants 128 --iters 1 > gpurun_out/ncu_wsb1.log 2>&1; tail -2 gpurun_out/ncu_wsb1.log

### run_r02d.sh
# round 2, call D: ws forward with 4 stages / lag 3, dual-stream consumer (2 CTAs/SM, 128 regs), with / without L2 prefetch
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "cooperative_and_per_ray or config3 or jitter") > gpurun_out/tests.log 2>&1; tail -4 gpurun_out/tests.log
timeout 300 python profiles/ab_kernels.py --variants 0,96,352,1120,1376,1376 --iters 10 > gpurun_out/ab_d.json 2> gpurun_out/ab_d.err; tail -7 gpurun_out/ab_d.err
timeout 300 python profiles/ab_kernels.py --variants 0,352,1376 --iters 10 --density-shift 0.9 > gpurun_out/ab_d_sparse.json 2> gpurun_out/ab_d_sparse.err; tail -4 gpurun_out/ab_d_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_ws -s 3 -c 1 -f -o gpurun_out/r02_ws3_dq python profiles/ab_kernels.py --variants 1376 --iters 1 > gpurun_out/ncu_ws3.log 2>&1; tail -2 gpurun_out/ncu_ws3.log

### run_r02e.sh
# round 2, call E: FFMA2 in the cooperative backward's member sweep; ws forward at 4 CTAs/SM (64 regs) vs 3
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libr3d_b200.so libr3d_b200_ws4.so libr3d_b200_ws3s3.so; do
  R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/$lib timeout 300 python profiles/ab_kernels.py --variants 0,352 --iters 10 > gpurun_out/ab_e_$lib.json 2> gpurun_out/ab_e_$lib.err
  tail -2 gpurun_out/ab_e_$lib.err
done
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training_parity.py tests/test_gpu_dual_render.py -m gpu -x -q) > gpurun_out/tests.log 2>&1; tail -4 gpurun_out/tests.log

### run_r02f.sh
# round 2, call F: lane-group forward with cell-sorted publish (hardware coalescing of same-cell gathers)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "cooperative_and_per_ray") > gpurun_out/tests.log 2>&1; tail -4 gpurun_out/tests.log
timeout 300 python profiles/ab_kernels.py --variants 0,2048,0,2048 --iters 10 > gpurun_out/ab_f.json 2> gpurun_out/ab_f.err; tail -5 gpurun_out/ab_f.err
timeout 300 python profiles/ab_kernels.py --variants 0,2048 --iters 10 --density-shift 0.9 > gpurun_out/ab_f_sparse.json 2> gpurun_out/ab_f_sparse.err; tail -3 gpurun_out/ab_f_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_group -s 3 -c 1 -f -o gpurun_out/r02_group_sort python profiles/ab_kernels.py --variants 2048 --iters 1 > gpurun_out/ncu_gs.log 2>&1; tail -2 gpurun_out/ncu_gs.log

### run_r02g.sh
# round 2, call G (2 GPUs): fused reduce-scatter/Adam/all-gather parity test, 1- and 2-GPU bench lines with the optimizer in the step
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q) > gpurun_out/tests_mg.log 2>&1; tail -5 gpurun_out/tests_mg.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 1500 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_fused.json 2> gpurun_out/r02_bench_n2_fused.err; tail -c 1200 gpurun_out/r02_bench_n2_fused.json; tail -3 gpurun_out/r02_bench_n2_fused.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl > gpurun_out/r02_bench_n2_nccl.json 2> gpurun_out/r02_bench_n2_nccl.err; tail -c 600 gpurun_out/r02_bench_n2_nccl.json; tail -3 gpurun_out/r02_bench_n2_nccl.err

### run_r02h.sh
# round 2, call H (2 GPUs): full GPU suite on the product build and on the A/B build, 2-GPU tests, traffic measurement, bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_prod.log 2>&1; tail -4 gpurun_out/tests_prod.log
(time R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/libr3d_b200_ab.so timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "cooperative_and_per_ray or contribution or mask") > gpurun_out/tests_ab.log 2>&1; tail -4 gpurun_out/tests_ab.log
timeout 600 python profiles/measure_traffic.py > gpurun_out/traffic.log 2>&1; tail -2 gpurun_out/traffic.log; cp profiles/traffic.json gpurun_out/traffic.json
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 700 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_fused.json 2> gpurun_out/r02_bench_n2_fused.err; tail -c 900 gpurun_out/r02_bench_n2_fused.json; tail -3 gpurun_out/r02_bench_n2_fused.err

### run_r02i.sh
# round 2, call I (8 GPUs): 2-GPU parity tests, bench at N = 8 / 4 (fused in-switch optimizer), N = 8 with NCCL, c5 at N = 8
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q) > gpurun_out/tests_mg.log 2>&1; tail -3 gpurun_out/tests_mg.log
run() { n=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
run 8 29521 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8_fused.json 2> gpurun_out/r02_bench_n8_fused.err; tail -c 400 gpurun_out/r02_bench_n8_fused.json; tail -2 gpurun_out/r02_bench_n8_fused.err
run 4 29522 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4_fused.json 2> gpurun_out/r02_bench_n4_fused.err; tail -c 400 gpurun_out/r02_bench_n4_fused.json; tail -2 gpurun_out/r02_bench_n4_fused.err
run 8 29523 --steps 10 --warmup 3 --exchange nccl > gpurun_out/r02_bench_n8_nccl.json 2> gpurun_out/r02_bench_n8_nccl.err; tail -c 300 gpurun_out/r02_bench_n8_nccl.json; tail -2 gpurun_out/r02_bench_n8_nccl.err
run 8 29524 --steps 10 --warmup 3 --exchange nvls > gpurun_out/r02_bench_n8_nvls_allreduce.json 2> gpurun_out/r02_bench_n8_nvls_allreduce.err; tail -c 300 gpurun_out/r02_bench_n8_nvls_allreduce.json; tail -2 gpurun_out/r02_bench_n8_nvls_allreduce.err
run 8 29525 --steps 5 --warmup 3 --workload c5_512cube_deg3_1600px_512spp > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err; tail -c 600 gpurun_out/r02_bench_c5_n8.json; tail -3 gpurun_out/r02_bench_c5_n8.err

### run_r02j.sh
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q) > gpurun_out/tests_mg.log 2>&1; grep -n "^E  " gpurun_out/tests_mg.log | tail -8; tail -2 gpurun_out/tests_mg.log
(time timeout 600 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "training_batch_sampler") > gpurun_out/tests_sampler.log 2>&1; tail -3 gpurun_out/tests_sampler.log

### run_r02k.sh
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q) > gpurun_out/tests_mg.log 2>&1; grep -n "^E  " gpurun_out/tests_mg.log | tail -5; tail -2 gpurun_out/tests_mg.log
run() { n=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
run 2 29541 --steps 10 --warmup 3 --strong > gpurun_out/r02_bench_n2_strong.json 2> gpurun_out/r02_bench_n2_strong.err; tail -c 300 gpurun_out/r02_bench_n2_strong.json; tail -2 gpurun_out/r02_bench_n2_strong.err
run 2 29542 --steps 10 --warmup 3 --exchange nvls > gpurun_out/r02_bench_n2_nvls.json 2> gpurun_out/r02_bench_n2_nvls.err; tail -c 300 gpurun_out/r02_bench_n2_nvls.json; tail -2 gpurun_out/r02_bench_n2_nvls.err
run 2 29543 --steps 10 --warmup 3 --no-optimizer > gpurun_out/r02_bench_n2_noopt.json 2> gpurun_out/r02_bench_n2_noopt.err; tail -c 300 gpurun_out/r02_bench_n2_noopt.json; tail -2 gpurun_out/r02_bench_n2_noopt.err

### run_r02l.sh
# round 2, call L: compute-sanitizer racecheck + memcheck on the c2 shape with in-kernel jitter (forward + backward, default kernels),
#                  2-GPU test re-run happens elsewhere
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python profiles/ab_kernels.py --variants 0 --grid 128 --side 400 --spp 128 --iters 1 --warmup 0 > gpurun_out/r02_racecheck_c2.log 2>&1; tail -6 gpurun_out/r02_racecheck_c2.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python profiles/ab_kernels.py --variants 0 --grid 128 --side 400 --spp 128 --iters 1 --warmup 0 > gpurun_out/r02_memcheck_c2.log 2>&1; tail -6 gpurun_out/r02_memcheck_c2.log
timeout 600 python profiles/ab_kernels.py --variants 0 --density-shift 0.9 > gpurun_out/r02_sparse.json 2> gpurun_out/r02_sparse.err; tail -2 gpurun_out/r02_sparse.err
timeout 600 python profiles/extra_bench.py c2 train > gpurun_out/r02_extra.json 2> gpurun_out/r02_extra.err; tail -c 800 gpurun_out/r02_extra.json
timeout 600 python profiles/dual_bench.py > gpurun_out/r02_dual.json 2> gpurun_out/r02_dual.err; tail -c 500 gpurun_out/r02_dual.json

### run_r02m.sh
# round 2, call M: lane-group forward with next-sample prefetch (L2 features; + L1 density)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/libr3d_b200_ab.so
timeout 300 python profiles/ab_kernels.py --variants 0,4096,8192,0,4096,8192 --iters 10 > gpurun_out/ab_m.json 2> gpurun_out/ab_m.err; tail -7 gpurun_out/ab_m.err
timeout 300 python profiles/ab_kernels.py --variants 0,4096,8192 --iters 10 --density-shift 0.9 > gpurun_out/ab_m_sparse.json 2> gpurun_out/ab_m_sparse.err; tail -4 gpurun_out/ab_m_sparse.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_group -s 3 -c 1 -f -o gpurun_out/r02_group_pf python profiles/ab_kernels.py --variants 4096 --iters 1 > gpurun_out/ncu_gpf.log 2>&1; tail -2 gpurun_out/ncu_gpf.log

### run_r02n.sh
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/libr3d_b200_ab_fb5.so
timeout 300 python profiles/ab_kernels.py --variants 0,4096,2048,0,4096,2048 --iters 10 > gpurun_out/ab_n.json 2> gpurun_out/ab_n.err; tail -7 gpurun_out/ab_n.err

### run_r02o.sh
# round 2, call O: lane-group forward with TMA density bricks (R3D_FWD_TMA=1) and TMA feature-brick L2 prefetch (=2)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for mode in 0 1 2; do
  R3D_FWD_TMA=$mode timeout 300 python profiles/ab_kernels.py --variants 0,0 --iters 10 > gpurun_out/ab_o_tma$mode.json 2> gpurun_out/ab_o_tma$mode.err; tail -2 gpurun_out/ab_o_tma$mode.err
  R3D_FWD_TMA=$mode timeout 300 python profiles/ab_kernels.py --variants 0 --iters 10 --density-shift 0.9 > gpurun_out/ab_o_tma${mode}_sparse.json 2> gpurun_out/ab_o_tma${mode}_sparse.err; tail -1 gpurun_out/ab_o_tma${mode}_sparse.err
done
(time R3D_FWD_TMA=2 timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_tma.log 2>&1; tail -4 gpurun_out/tests_tma.log
R3D_FWD_TMA=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_group_tma -s 3 -c 1 -f -o gpurun_out/r02_group_tma python profiles/ab_kernels.py --variants 0 --iters 1 > gpurun_out/ncu_gtma.log 2>&1; tail -2 gpurun_out/ncu_gtma.log

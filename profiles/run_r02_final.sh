# round 2, final evidence call (1 GPU): GPU suite, smoke, traffic measurement (digest-stamped), driver-style bench line + reference arm,
# launch list of the bench command, ncu --set full of the two product kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_final.log 2>&1 < /dev/null; tail -3 gpurun_out/tests_final.log
timeout 300 python -m pytest tests/test_gpu_training_parity.py -m gpu -q -s -k trainer_like < /dev/null 2>&1 | grep "held-out" > gpurun_out/psnr_parity.log; cat gpurun_out/psnr_parity.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1 < /dev/null; tail -2 gpurun_out/smoke_final.log
timeout 600 python profiles/measure_traffic.py > gpurun_out/traffic.log 2>&1 < /dev/null; tail -1 gpurun_out/traffic.log | cut -c1-200; cp profiles/traffic.json gpurun_out/traffic.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err < /dev/null; tail -c 300 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err < /dev/null; tail -c 300 gpurun_out/r02_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 < /dev/null > gpurun_out/launches_run.log 2>&1; grep -c r3d gpurun_out/r02_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_fwd_group -s 3 -c 1 -f -o gpurun_out/r02_final_fwd python profiles/ab_kernels.py --variants 0 --iters 1 < /dev/null > gpurun_out/ncu_final_fwd.log 2>&1; tail -1 gpurun_out/ncu_final_fwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_bwd_coop -s 3 -c 1 -f -o gpurun_out/r02_final_bwd python profiles/ab_kernels.py --variants 0 --iters 1 < /dev/null > gpurun_out/ncu_final_bwd.log 2>&1; tail -1 gpurun_out/ncu_final_bwd.log
timeout 300 python profiles/ab_kernels.py --variants 0 --iters 10 --density-shift 0.9 < /dev/null > gpurun_out/r02_sparse.json 2> gpurun_out/r02_sparse.err; tail -1 gpurun_out/r02_sparse.err | cut -c1-200

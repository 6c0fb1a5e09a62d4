# round 2, final evidence call (1 GPU): GPU suite, traffic measurement, driver-style bench line, launch list, ncu --set full of both kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_final.log 2>&1; tail -3 gpurun_out/tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
timeout 600 python profiles/measure_traffic.py > gpurun_out/traffic.log 2>&1; tail -1 gpurun_out/traffic.log | cut -c1-200; cp profiles/traffic.json gpurun_out/traffic.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 500 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 400 gpurun_out/r02_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"render_(fwd|bwd)" -s 6 -c 2 -f -o gpurun_out/r02_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02_full.log 2>&1; tail -2 gpurun_out/r02_full.log

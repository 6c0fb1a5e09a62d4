# round 2, final evidence call (1 GPU): GPU suite, smoke, traffic measurement (digest-stamped), driver-style bench line + reference arm
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_final.log 2>&1 < /dev/null; tail -3 gpurun_out/tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1 < /dev/null; tail -2 gpurun_out/smoke_final.log
timeout 600 python profiles/measure_traffic.py > gpurun_out/traffic.log 2>&1 < /dev/null; tail -1 gpurun_out/traffic.log | cut -c1-200; cp profiles/traffic.json gpurun_out/traffic.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err < /dev/null; tail -c 300 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err < /dev/null; tail -c 300 gpurun_out/r02_bench_reference.json

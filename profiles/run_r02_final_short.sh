# round 2, short evidence call (1 GPU) after a change that does not touch the render kernels: GPU suite, traffic (digest-stamped), bench line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/tests_final.log 2>&1 < /dev/null; grep -E "passed|failed" gpurun_out/tests_final.log
timeout 600 python profiles/measure_traffic.py > gpurun_out/traffic.log 2>&1 < /dev/null; tail -1 gpurun_out/traffic.log | cut -c1-120; cp profiles/traffic.json gpurun_out/traffic.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err < /dev/null; tail -2 gpurun_out/r02_bench_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read())
print(round(d['value']/1e6,2), round(d['ms_per_step'],3), round(d['e2e']['value']/1e6,2), {k:round(v,3) for k,v in d['phases_ms'].items()}, d['roofline']['traffic'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 < /dev/null > gpurun_out/launches_run.log 2>&1; grep -c r3d gpurun_out/r02_launches.csv

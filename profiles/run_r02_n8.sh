# round 2, multi-GPU call (N = $1): driver-style bench line with the final library
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 < /dev/null > gpurun_out/r02_bench_n${N}_fused.json 2> gpurun_out/r02_bench_n${N}_fused.err
python - <<PY
import json
txt=open('gpurun_out/r02_bench_n${N}_fused.json').read()
line=[l for l in txt.splitlines() if l.startswith('{')][-1]
d=json.loads(line); print(len(txt.splitlines()), round(d['value']/1e6,2), round(d['ms_per_step'],3), round(d['e2e']['value']/1e6,2), {k:round(v,3) for k,v in d['phases_ms'].items()}, d['parity_check']['ok'], d['config']['exchange'], d['clocks'])
PY
tail -2 gpurun_out/r02_bench_n${N}_fused.err

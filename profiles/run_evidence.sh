# usage (GPU box): bash profiles/run_evidence.sh <tag>   -- bench line, ncu launch list, ncu --set full of both render kernels
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
tag=$1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 600 gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"render_(fwd|bwd)" -s 6 -c 2 -f -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_full.log 2>&1
tail -2 gpurun_out/${tag}_full.log
python profiles/ab_kernels.py --variants 0 --density-shift 0.9 > gpurun_out/${tag}_sparse.json 2> gpurun_out/${tag}_sparse.err
tail -2 gpurun_out/${tag}_sparse.err
python profiles/extra_bench.py c2 train > gpurun_out/${tag}_extra.json 2> gpurun_out/${tag}_extra.err
python profiles/dual_bench.py > gpurun_out/${tag}_dual.json 2> gpurun_out/${tag}_dual.err
tail -c 1500 gpurun_out/${tag}_extra.json

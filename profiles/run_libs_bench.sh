# usage (GPU box): bash profiles/run_libs_bench.sh lib1.so lib2.so ...   -- bench phases with differently compiled libraries
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
for lib in "$@"; do
  R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline < /dev/null 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['phases_ms'].items()})"
done

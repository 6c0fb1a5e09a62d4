# usage (on the GPU box): MODES="0 1 2" [CARVES="-1 50"] bash profiles/run_split3.sh  -- two-kernel forward, gather modes (measurement build)
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
export R3D_LIB_PATH=$GRAFT_REPO_ROOT/thr3ed_atom_b200/_lib/libr3d_b200_ab.so
for m in ${MODES:-0 1 2}; do
 for cv in ${CARVES:--1}; do
  for shift in ${SHIFTS:-0.0 0.9}; do
    echo "mode $m carve $cv shift $shift"
    R3D_SPLIT_CARVEOUT=$cv R3D_SPLIT_MODE=$m timeout 300 python profiles/ab_kernels.py --variants ${VARIANTS:-32768} --iters 10 --density-shift $shift < /dev/null 2>&1 > /dev/null | grep variant | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: (round(v, 4) if isinstance(v, float) and k.endswith('ms') else v) for k, v in d.items() if k in ('variant', 'fwd_ms', 'bwd_ms', 'step_ms', 'colour_sum')})
"
  done
 done
done 2>&1 | tee gpurun_out/split_modes.log

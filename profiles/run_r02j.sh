set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q) > gpurun_out/tests_mg.log 2>&1; grep -n "^E  " gpurun_out/tests_mg.log | tail -8; tail -2 gpurun_out/tests_mg.log
(time timeout 600 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "training_batch_sampler") > gpurun_out/tests_sampler.log 2>&1; tail -3 gpurun_out/tests_sampler.log

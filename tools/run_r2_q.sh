cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/mc_replays.py 2>&1 | tee gpurun_out/r2q_mc_replays.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_tests.txt 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2q_tests.txt

cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_philox.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
python tools/mc_speed.py

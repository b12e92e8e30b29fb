cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_multipath.py -m gpu -q -x 2>&1 | grep -E "^E|assert|Error" | head -20

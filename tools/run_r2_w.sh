cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -2
for k in tx_fast tx_exact; do timeout 300 python tools/r2_kernels.py $k 10 2>&1 | tail -1; done

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_fullpath.py tests/test_ref_compat.py -m gpu -q 2>&1 | tail -4
timeout 300 python tools/fft_probe.py 8 2>&1 | tee gpurun_out/r2s_fft_probe.txt
timeout 600 python tools/stage_probe.py 2 2>&1 | tee gpurun_out/r2s_stage_probe.txt

# round 2, call Y: k_stream_quad NSYM2 build (compile-time ring slots): receiver tests, timings against the general build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stream_quad.py tests/test_gpu_checked.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_sweep_lin.py tests/test_gpu_multipath.py -m gpu -q -x 2>&1 | tail -4
for k in rx_fast rx_exact point point_fast; do echo -n "nsym2 build    "; timeout 120 python tools/r2_kernels.py $k 20 2>&1 | tail -1; done | tee gpurun_out/r2y_nsym2.txt

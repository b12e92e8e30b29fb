# round 2, call L: k_stream_quad, 6 vs 8 warps per block vs the one-frame-per-warp kernels; parity subset; ncu of the 6-warp build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_checked.py tests/test_gpu_parity.py tests/test_gpu_long_frames.py tests/test_gpu_golden.py tests/test_gpu_multipath.py -m gpu -q > gpurun_out/r2l_tests.txt 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2l_tests.txt
for cfg in "0 6" "0 8" "1 8"; do
  set -- $cfg
  for k in rx_fast rx_exact point point_fast; do
    echo -n "layout $1 warps $2  "; STREAM_LAYOUT=$1 STREAM_WARPS=$2 timeout 300 python tools/r2_kernels.py $k 20 2>&1 | tail -1
  done
done | tee gpurun_out/r2l_ab.txt
for what in rx_fast rx_exact point; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_quad -s 1 -c 1 -f -o gpurun_out/r2l_prof_$what python tools/r2_kernels.py $what 2 > gpurun_out/r2l_ncu_$what.log 2>&1
  echo "ncu $what rc=$?"
  python tools/ncu_summary.py gpurun_out/r2l_prof_$what.ncu-rep gpurun_out/r2l_ncu_$what.txt > /dev/null 2>&1
done

# round 2, call K: k_stream_quad iteration: quick parity subset, A/B timing of the two streaming layouts, ncu of the new kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_checked.py tests/test_gpu_parity.py tests/test_gpu_long_frames.py tests/test_gpu_golden.py -m gpu -q > gpurun_out/r2k_tests.txt 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2k_tests.txt
for lay in 0 1; do
  for k in rx_fast rx_exact point point_fast; do
    echo -n "layout $lay  "; STREAM_LAYOUT=$lay timeout 300 python tools/r2_kernels.py $k 20 2>&1 | tail -1
  done
done | tee gpurun_out/r2k_ab.txt
for what in rx_fast point; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_quad -s 1 -c 1 -f -o gpurun_out/r2k_prof_$what python tools/r2_kernels.py $what 2 > gpurun_out/r2k_ncu_$what.log 2>&1
  echo "ncu $what rc=$?"
  python tools/ncu_summary.py gpurun_out/r2k_prof_$what.ncu-rep gpurun_out/r2k_ncu_$what.txt > /dev/null 2>&1
done

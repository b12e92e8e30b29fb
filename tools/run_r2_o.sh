# round 2, call O: k_mc_quad (fused Monte-Carlo, one frame per lane group): GPU suite, timings of both layouts
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_tests.txt 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r2o_tests.txt
for lay in 0 1; do
  for k in mc_fast mc_exact; do
    echo -n "layout $lay  "; STREAM_LAYOUT=$lay timeout 300 python tools/r2_kernels.py $k 6 2>&1 | tail -1
  done
done | tee gpurun_out/r2o_ab.txt

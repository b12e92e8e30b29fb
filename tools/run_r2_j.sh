# round 2, call J: first run of k_stream_quad (one frame per lane group): GPU suite, then A/B of the two streaming layouts on one box
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.txt 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2j_tests.txt
for rep in 1 2; do
  for lay in 0 1; do
    for k in rx_fast rx_exact point point_fast; do
      echo -n "layout $lay  "; STREAM_LAYOUT=$lay timeout 300 python tools/r2_kernels.py $k 20 2>&1 | tail -1
    done
  done
done | tee gpurun_out/r2j_ab.txt

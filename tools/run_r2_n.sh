# round 2, call N: ring depth of k_stream_quad (slots per warp without / with draws): 4/3 (default) vs 6/4 vs 3/2, alternating on one box
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
  for v in d43 d64 d32; do
    for k in rx_fast rx_exact point; do
      echo -n "$v  "; OFDM_B200_LIB=$GRAFT_REPO_ROOT/build/ab/lib_$v.so timeout 300 python tools/r2_kernels.py $k 20 2>&1 | tail -1
    done
  done
done | tee gpurun_out/r2n_depth_ab.txt

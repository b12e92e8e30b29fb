# round 2, call M: k_stream_quad with per-frame thresholds: full GPU suite, soak (checked == all-exact totals), timings
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2m_tests.txt 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2m_tests.txt
for k in rx_fast rx_exact point point_fast; do
  echo -n "quad, 6 warps  "; timeout 300 python tools/r2_kernels.py $k 20 2>&1 | tail -1
done | tee gpurun_out/r2m_ab.txt
timeout 1200 python tools/checked_soak.py 2 3 > gpurun_out/r2m_soak.txt 2>&1; echo "soak rc=$?"; grep -c "equal True" gpurun_out/r2m_soak.txt; grep "equal False" gpurun_out/r2m_soak.txt | head; grep -E "decisions|multipath|n_sym|Monte" gpurun_out/r2m_soak.txt | tail -22

# round 2, call U: k_mc_quad<fast, multipath>: multipath tests, timings of both layouts, ncu summary
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multipath.py tests/test_gpu_philox.py tests/test_gpu_checked.py -m gpu -q 2>&1 | tail -8
for lay in 0 1; do echo -n "layout $lay  "; STREAM_LAYOUT=$lay timeout 300 python tools/r2_kernels.py mp_fast 6 2>&1 | tail -1; done | tee gpurun_out/r2u_mp.txt
timeout 600 ncu --set full --clock-control none -k regex:k_mc_quad -s 1 -c 1 -f -o gpurun_out/r2u_prof_mp_fast python tools/r2_kernels.py mp_fast 2 > gpurun_out/r2u_ncu_mp_fast.log 2>&1
python tools/ncu_summary.py gpurun_out/r2u_prof_mp_fast.ncu-rep gpurun_out/r2u_ncu_mp_fast.txt > /dev/null 2>&1; grep -E "k_mc_quad|time_duration|inst_executed.sum|issue_active|per_cycle|stall|registers" gpurun_out/r2u_ncu_mp_fast.txt
rm -f gpurun_out/r2u_prof_mp_fast.ncu-rep

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/mc_replays.py 2>&1 | tee gpurun_out/r2p_mc_replays.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mc_quad -s 1 -c 1 -f -o gpurun_out/r2p_prof_mc_exact python tools/r2_kernels.py mc_exact 2 > gpurun_out/r2p_ncu_mc_exact.log 2>&1
python tools/ncu_summary.py gpurun_out/r2p_prof_mc_exact.ncu-rep gpurun_out/r2p_ncu_mc_exact.txt > /dev/null 2>&1; grep -E "time_duration|inst_executed.sum|issue_active|per_cycle|stall|registers" gpurun_out/r2p_ncu_mc_exact.txt
timeout 600 ncu --set full --clock-control none -k regex:k_mc_quad -s 1 -c 1 -f -o gpurun_out/r2p_prof_mc_fast python tools/r2_kernels.py mc_fast 2 > gpurun_out/r2p_ncu_mc_fast.log 2>&1
python tools/ncu_summary.py gpurun_out/r2p_prof_mc_fast.ncu-rep gpurun_out/r2p_ncu_mc_fast.txt > /dev/null 2>&1; grep -E "time_duration|inst_executed.sum|issue_active|per_cycle|stall|registers" gpurun_out/r2p_ncu_mc_fast.txt
rm -f gpurun_out/r2p_prof_mc_fast.ncu-rep

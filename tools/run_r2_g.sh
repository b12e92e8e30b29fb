# round 2, call G: ncu --set full (with source) of the sweep kernel and the tiled power kernel after their plain runs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for k in sweep:k_sweep_lin power_full:k_frame_power_tiled; do
  what=${k%%:*}; kern=${k##*:}
  timeout 300 python tools/r2_kernels.py $what 2 > gpurun_out/r2g_plain_$what.log 2>&1 || { echo "plain $what failed"; continue; }
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -f -o gpurun_out/r2g_prof_$what python tools/r2_kernels.py $what 2 > gpurun_out/r2g_ncu_$what.log 2>&1
  echo "ncu $what rc=$?"
  python tools/ncu_summary.py gpurun_out/r2g_prof_$what.ncu-rep gpurun_out/r2g_ncu_$what.txt > /dev/null 2>&1
done

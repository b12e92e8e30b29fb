# round 2, call C: whole GPU suite, per-kernel timings, ncu captures (each only after its plain run exited 0), bench launch list
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.txt
tail -12 gpurun_out/r2c_pytest.txt
timeout 600 python tools/r2_kernels.py all 5 > gpurun_out/r2c_kernels.txt 2>&1; echo "kernels exit $?"; cat gpurun_out/r2c_kernels.txt
for k in sweep:k_sweep_lin point:k_stream_rx2 rx_fast:k_stream_rx2 rx_exact:k_stream_rx2 tx_fast:k_tx_frames2 mc_fast:k_mc_philox mc_exact:k_mc_philox; do
  what=${k%%:*}; kern=${k##*:}
  timeout 300 python tools/r2_kernels.py $what 2 > gpurun_out/r2c_plain_$what.log 2>&1 || { echo "plain $what failed"; continue; }
  src=""; [ "$what" = "sweep" ] && src="--import-source on"; [ "$what" = "point" ] && src="--import-source on"
  timeout 600 ncu --set full --clock-control none $src -k regex:$kern -s 1 -c 1 -f -o gpurun_out/r2c_prof_$what python tools/r2_kernels.py $what 2 > gpurun_out/r2c_ncu_$what.log 2>&1
  echo "ncu $what rc=$?"
  python tools/ncu_summary.py gpurun_out/r2c_prof_$what.ncu-rep gpurun_out/r2c_ncu_$what.txt > /dev/null 2>&1
  [ -z "$src" ] && rm -f gpurun_out/r2c_prof_$what.ncu-rep
done
python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r2c_plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r2c_ncu_bench.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out | grep r2c

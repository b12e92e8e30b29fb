# round 2, call H (final library of round 2): suite, per-kernel timings, ncu --set full of every kernel of interest (each after its plain run exited 0),
# bench launch list, default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export OFDM_TEST_LOG=$GRAFT_REPO_ROOT/gpurun_out/r2h_test_log.txt; rm -f $OFDM_TEST_LOG
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest.txt
tail -6 gpurun_out/r2h_pytest.txt
timeout 600 python tools/r2_kernels.py all 5 > gpurun_out/r2h_kernels.txt 2>&1; echo "kernels exit $?"; cat gpurun_out/r2h_kernels.txt
for k in power_full:k_frame_power_tiled sweep:k_sweep_lin sweep_fast:k_sweep_lin point:k_stream_rx2 point_fast:k_stream_rx2 rx_fast:k_stream_rx2 rx_exact:k_stream_rx2 tx_fast:k_tx_frames2 tx_exact:k_tx_frames2 mc_fast:k_mc_philox mc_exact:k_mc_philox mp_fast:k_mc_philox; do
  what=${k%%:*}; kern=${k##*:}
  timeout 300 python tools/r2_kernels.py $what 2 > gpurun_out/r2h_plain_$what.log 2>&1 || { echo "plain $what failed"; continue; }
  src=""; [ "$what" = "sweep" ] && src="--import-source on"; [ "$what" = "rx_fast" ] && src="--import-source on"
  timeout 600 ncu --set full --clock-control none $src -k regex:$kern -s 1 -c 1 -f -o gpurun_out/r2h_prof_$what python tools/r2_kernels.py $what 2 > gpurun_out/r2h_ncu_$what.log 2>&1
  echo "ncu $what rc=$?"
  python tools/ncu_summary.py gpurun_out/r2h_prof_$what.ncu-rep gpurun_out/r2h_ncu_$what.txt > /dev/null 2>&1
  [ -z "$src" ] && rm -f gpurun_out/r2h_prof_$what.ncu-rep
done
python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r2h_plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/r2h_ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err; echo "reference arm exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2h_bench.json'))
print('value %.3e e2e %.3e ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
r = d['roofline']; print('roofline sustained', round(r['frac'], 3), r['kernel_ms'], 'burst', round(r['frac_burst'], 3), r['kernel_ms_burst'])
print('sweep_kernel_ms', d['sweep_kernel']['kernel_ms'], 'cpu', d.get('cpu_baseline'))
print(json.load(open('gpurun_out/r2h_bench_reference.json'))['value'])
PY

# round 2, call D: suite on the FAST EVM guard / deferred reductions, per-kernel timings (+ blocks-per-SM A/B), default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export OFDM_TEST_LOG=$GRAFT_REPO_ROOT/gpurun_out/r2d_test_log.txt; rm -f $OFDM_TEST_LOG
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.txt
tail -12 gpurun_out/r2d_pytest.txt; cat $OFDM_TEST_LOG
timeout 600 python tools/r2_kernels.py all 5 > gpurun_out/r2d_kernels.txt 2>&1; echo "kernels exit $?"; cat gpurun_out/r2d_kernels.txt
for v in build/ab/lib_b2.so build/ab/lib_b3.so build/ab/lib_b2.so build/ab/lib_b3.so; do
  for k in rx_fast rx_exact; do echo -n "$v "; OFDM_B200_LIB=$GRAFT_REPO_ROOT/$v timeout 300 python tools/r2_kernels.py $k 5; done
done 2>&1 | tee gpurun_out/r2d_blocks_ab.txt
timeout 900 python bench.py --steps 10 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2d_bench.json'))
    print('value %.3e e2e %.3e ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
    print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'sweep_kernel_ms', d['sweep_kernel']['kernel_ms'], 'replayed', d['sweep_kernel']['points_replayed_exactly_per_sweep'])
    c = d['configs']
    for m in ('fast', 'exact'):
        print('cfg2', m, 'tx', round(c['cfg2_streaming'][m]['tx']['roofline']['frac'], 3), 'rx', round(c['cfg2_streaming'][m]['rx']['roofline']['frac'], 3))
        print('cfg3', m, '%.3e' % c['cfg3_philox_mc'][m]['symbols_per_s'], 'cfg4', '%.3e' % c['cfg4_multipath_8taps'][m]['symbols_per_s'])
    u = c['cfg3_philox_mc']['until_100_errors_or_1e-7']
    print('until', u['seconds'], u['rounds'])
except Exception as e:
    print('bench parse failed', e)
PY

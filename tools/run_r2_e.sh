# round 2, call E: suite, EVM-guard trade-off, per-kernel timings, default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export OFDM_TEST_LOG=$GRAFT_REPO_ROOT/gpurun_out/r2e_test_log.txt; rm -f $OFDM_TEST_LOG
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_pytest.txt
tail -12 gpurun_out/r2e_pytest.txt
timeout 600 python tools/guard_probe.py 1000000 > gpurun_out/r2e_guard_probe.txt 2>&1; cat gpurun_out/r2e_guard_probe.txt
timeout 600 python tools/r2_kernels.py all 5 > gpurun_out/r2e_kernels.txt 2>&1; echo "kernels exit $?"; cat gpurun_out/r2e_kernels.txt
timeout 900 python bench.py --steps 10 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r2e_bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2e_bench.json'))
    print('value %.3e e2e %.3e ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
    print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['kernel_ms_by_snr_point'])
    print('sweep_kernel_ms', d['sweep_kernel']['kernel_ms'], 'replayed', d['sweep_kernel']['points_replayed_exactly_per_sweep'])
    c = d['configs']
    for m in ('fast', 'exact'):
        print('cfg2', m, 'tx', round(c['cfg2_streaming'][m]['tx']['roofline']['frac'], 3), 'rx', round(c['cfg2_streaming'][m]['rx']['roofline']['frac'], 3))
        print('cfg3', m, '%.3e' % c['cfg3_philox_mc'][m]['symbols_per_s'], 'cfg4', '%.3e' % c['cfg4_multipath_8taps'][m]['symbols_per_s'])
except Exception as e:
    print('bench parse failed', e)
PY

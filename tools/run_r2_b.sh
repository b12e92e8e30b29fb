# round 2, call B: the whole GPU suite on the new kernels (k_sweep_lin, stage exports, until rule), then the default bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.txt
tail -15 gpurun_out/r2b_pytest.txt
timeout 900 python bench.py --steps 10 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2b_bench.json'))
    print('value %.3e e2e %.3e ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
    print('roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'sweep_kernel_ms', d['sweep_kernel']['kernel_ms'], 'replayed', d['sweep_kernel']['points_replayed_exactly_per_sweep'])
    c = d['configs']
    for m in ('fast', 'exact'):
        print('cfg2', m, 'tx', round(c['cfg2_streaming'][m]['tx']['roofline']['frac'], 3), 'rx', round(c['cfg2_streaming'][m]['rx']['roofline']['frac'], 3))
        print('cfg3', m, '%.3e' % c['cfg3_philox_mc'][m]['symbols_per_s'], 'cfg4', '%.3e' % c['cfg4_multipath_8taps'][m]['symbols_per_s'])
    u = c['cfg3_philox_mc']['until_100_errors_or_1e-7']
    print('until', u['seconds'], u['rounds'], [(p['snr_db'], p['bit_errors'], p['bits']) for p in u['points'][10:]])
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench parse failed', e)
PY

# final check of the shipped library: smoke(), the whole GPU suite, the default bench line, gather timing (extras)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_final_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2b_final_pytest.txt
timeout 900 python bench.py --extras > gpurun_out/r2b_final_bench_extras.json 2> gpurun_out/r2b_final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2b_final_bench_extras.json'))
print('value %.3e e2e %.3e roofline %.3f burst %.3f launches %d' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_burst'], d['gpu_launches']))
c = d['configs']
print('cfg2', {m: (round(c['cfg2_streaming'][m]['tx']['roofline']['frac'], 3), round(c['cfg2_streaming'][m]['rx']['roofline']['frac'], 3)) for m in ('fast', 'exact')})
print('cfg3', ['%.3e' % c['cfg3_philox_mc'][m]['symbols_per_s'] for m in ('fast', 'exact')], 'cfg4', ['%.3e' % c['cfg4_multipath_8taps'][m]['symbols_per_s'] for m in ('fast', 'exact')])
ex = d['extras']['next_rows_full_receiver_path']
for s in ex['stages']: print('  %-48s %.3f ms  %.2f' % (s['stage'], s['ms'], s['frac_of_hbm_peak']))
print('full path', ex['total_ms'], 'ms', '%.3e frames/s' % ex['frames_per_s'])
print(d['clocks'])
PY

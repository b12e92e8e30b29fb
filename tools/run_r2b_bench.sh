# the default bench line of the shipped library (profiles/r2b_bench.json) + transmitter timings
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
for k in tx_fast tx_exact; do timeout 300 python tools/r2_kernels.py $k 10 2>&1 | tail -1; done | tee gpurun_out/r2b_tx_timings.txt
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2b_bench.json'))
print('value %.3e e2e %.3e ms/step %.3f roofline %.3f burst %.3f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_burst'], d['gpu_launches']))
c = d['configs']
print('cfg2', {m: (round(c['cfg2_streaming'][m]['tx']['roofline']['frac'], 3), round(c['cfg2_streaming'][m]['rx']['roofline']['frac'], 3)) for m in ('fast', 'exact')})
print('cfg3', ['%.3e' % c['cfg3_philox_mc'][m]['symbols_per_s'] for m in ('fast', 'exact')], 'cfg4', ['%.3e' % c['cfg4_multipath_8taps'][m]['symbols_per_s'] for m in ('fast', 'exact')], c['cfg4_multipath_8taps']['fast']['roofline'].get('warp_instructions_per_unit'))
print(d['clocks'], d['cpu_baseline']['value'])
PY

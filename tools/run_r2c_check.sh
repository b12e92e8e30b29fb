# last check of HEAD: smoke(), the whole GPU suite, default bench line, reference arm
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_check_pytest.txt 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2c_check_pytest.txt
timeout 900 python bench.py > gpurun_out/r2c_check_bench.json 2> gpurun_out/r2c_check_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c_check_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_check_reference.json 2>/dev/null; echo "reference rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c_check_bench.json'))
r = d['roofline']
print('value %.3e e2e %.3e' % (d['value'], d['e2e']['value']), 'roofline', round(r['frac'], 3), round(r['frac_burst'], 3), 'of nominal', round(r['frac_of_nominal'], 3), 'cpu', d['cpu_baseline']['value'])
c2 = d['configs']['cfg2_streaming']
print('cfg2', {m: (round(c2[m]['tx']['roofline']['frac'], 3), round(c2[m]['rx']['roofline']['frac'], 3)) for m in ('fast', 'exact')})
print('reference arm', json.load(open('gpurun_out/r2c_check_reference.json'))['value'])
PY

# last check of HEAD: smoke(), default bench line (with the CPU baseline), reference arm
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2c_check_bench.json 2> gpurun_out/r2c_check_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c_check_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_check_reference.json 2>/dev/null; echo "reference rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c_check_bench.json'))
r = d['roofline']
print('value %.3e e2e %.3e' % (d['value'], d['e2e']['value']), 'roofline', round(r['frac'], 3), round(r['frac_burst'], 3), 'of nominal', round(r['frac_of_nominal'], 3), 'cpu', d['cpu_baseline']['value'])
print('reference arm', json.load(open('gpurun_out/r2c_check_reference.json'))['value'])
PY

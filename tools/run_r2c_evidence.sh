# round 2, last evidence pass (k_stream_quad with compile-time ring slots for two-symbol frames): GPU suite, kernel timings, ncu of the four
# k_stream_quad variants, bench launch list, default bench line, soak
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r2c_
timeout 1500 python -m pytest tests -m gpu -q > ${P}pytest.txt 2>&1; echo "pytest exit $?" >> ${P}pytest.txt; tail -3 ${P}pytest.txt
timeout 600 python tools/r2_kernels.py all 5 > ${P}kernels.txt 2>&1; echo "kernels exit $?"; cat ${P}kernels.txt
for what in point point_fast rx_fast rx_exact; do
  timeout 300 python tools/r2_kernels.py $what 2 > ${P}plain_$what.log 2>&1 || { echo "plain $what failed"; continue; }
  timeout 600 ncu --set full --clock-control none -k regex:k_stream_quad -s 1 -c 1 -f -o ${P}prof_$what python tools/r2_kernels.py $what 2 > ${P}ncu_$what.log 2>&1
  echo "ncu $what rc=$?"
  python tools/ncu_summary.py ${P}prof_$what.ncu-rep ${P}ncu_$what.txt > /dev/null 2>&1
  rm -f ${P}prof_$what.ncu-rep
done
python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > ${P}plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > ${P}ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 python bench.py > ${P}bench.json 2> ${P}bench.err; echo "bench exit $?"
timeout 900 python tools/checked_soak.py 4 6 > ${P}soak.txt 2>&1; echo "soak rc=$?"; grep -c "equal True" ${P}soak.txt; grep -c "equal False" ${P}soak.txt; grep decisions ${P}soak.txt
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c_bench.json'))
print('value %.3e e2e %.3e ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
r = d['roofline']; print('roofline', r['kernel'], 'sustained', round(r['frac'], 3), r['kernel_ms'], 'burst', round(r['frac_burst'], 3), r['kernel_ms_burst'])
c2 = d['configs']['cfg2_streaming']
for m in ('fast', 'exact'):
    print('cfg2', m, 'tx', round(c2[m]['tx']['roofline']['frac'], 3), 'rx', round(c2[m]['rx']['roofline']['frac'], 3))
c = d['configs']
print('cfg3', ['%.3e' % c['cfg3_philox_mc'][m]['symbols_per_s'] for m in ('fast', 'exact')], 'cfg4', ['%.3e' % c['cfg4_multipath_8taps'][m]['symbols_per_s'] for m in ('fast', 'exact')])
PY

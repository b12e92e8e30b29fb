# round 2, second half (k_stream_quad): evidence for profiles/r2b_*: GPU suite, per-kernel timings (both streaming layouts), ncu --set full of the
# four k_stream_quad and two k_mc_quad variants (each after its plain run exited 0), bench launch list, default bench line, reference arm, soak, full-size parity
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
P=gpurun_out/r2b_
timeout 1500 python -m pytest tests -m gpu -q > ${P}pytest.txt 2>&1; echo "pytest exit $?" >> ${P}pytest.txt; tail -3 ${P}pytest.txt
timeout 600 python tools/r2_kernels.py all 5 > ${P}kernels.txt 2>&1; echo "kernels exit $?"; cat ${P}kernels.txt
for k in point point_fast rx_fast rx_exact mc_fast mc_exact mp_fast; do echo -n "one frame per warp (stream_layout 1)  "; STREAM_LAYOUT=1 timeout 300 python tools/r2_kernels.py $k 5 2>&1 | tail -1; done | tee ${P}kernels_layout1.txt
for k in point rx_fast rx_exact; do echo -n "8 warps per block (stream_warps 8)  "; STREAM_WARPS=8 timeout 300 python tools/r2_kernels.py $k 5 2>&1 | tail -1; done | tee ${P}kernels_warps8.txt
timeout 600 python tools/nsym_probe.py > ${P}nsym_probe.txt 2>&1; tail -12 ${P}nsym_probe.txt
for what in point point_fast rx_fast rx_exact mc_fast mc_exact mp_fast; do
  timeout 300 python tools/r2_kernels.py $what 2 > ${P}plain_$what.log 2>&1 || { echo "plain $what failed"; continue; }
  src=""; [ "$what" = "rx_exact" ] && src="--import-source on"
  timeout 600 ncu --set full --clock-control none $src -k 'regex:k_stream_quad|k_mc_quad' -s 1 -c 1 -f -o ${P}prof_$what python tools/r2_kernels.py $what 2 > ${P}ncu_$what.log 2>&1
  echo "ncu $what rc=$?"
  python tools/ncu_summary.py ${P}prof_$what.ncu-rep ${P}ncu_$what.txt > /dev/null 2>&1
  [ -z "$src" ] && rm -f ${P}prof_$what.ncu-rep
done
python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > ${P}plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > ${P}ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 900 python bench.py > ${P}bench.json 2> ${P}bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > ${P}bench_reference.json 2> ${P}bench_reference.err; echo "reference arm exit $?"
timeout 900 python tools/checked_soak.py 4 6 > ${P}soak.txt 2>&1; echo "soak rc=$?"; grep -c "equal True" ${P}soak.txt; grep -c "equal False" ${P}soak.txt; grep decisions ${P}soak.txt
timeout 600 python tools/full_parity.py --out ${P}full_parity.json > ${P}full_parity.txt 2>&1; echo "full parity rc=$?"; tail -2 ${P}full_parity.txt | cut -c1-300
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2b_bench.json'))
print('value %.3e e2e %.3e ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
r = d['roofline']; print('roofline', r['kernel'], 'sustained', round(r['frac'], 3), r['kernel_ms'], 'burst', round(r['frac_burst'], 3), r['kernel_ms_burst'])
c2 = d['configs']['cfg2_streaming']
for m in ('fast', 'exact'):
    print('cfg2', m, 'tx', round(c2[m]['tx']['roofline']['frac'], 3), 'rx', round(c2[m]['rx']['roofline']['frac'], 3))
print('sweep_kernel_ms', d['sweep_kernel']['kernel_ms'], 'cpu', d.get('cpu_baseline'))
print(json.load(open('gpurun_out/r2b_bench_reference.json'))['value'])
PY

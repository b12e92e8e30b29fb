# A/B two builds of libofdm_b200.so on the same box (per-launch time of the sweep kernel): build/ab/lib_a.so vs lib_b.so
cd $GRAFT_REPO_ROOT
for rep in 1 2 3; do
  for v in a b; do
    OFDM_B200_LIB=$GRAFT_REPO_ROOT/build/ab/lib_$v.so python bench.py --no-cpu --steps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['roofline']['kernel_ms'],4), round(d['value']/1e9,3), round(d['e2e']['value']/1e9,3))"
  done
done

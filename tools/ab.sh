# A/B of library builds on the same box: tools/ab.sh <lib> <lib> ...   (paths relative to the repo root, e.g. build/ab/lib_a.so)
cd $GRAFT_REPO_ROOT
for rep in 1 2 3; do
  for v in "$@"; do
    OFDM_B200_LIB=$GRAFT_REPO_ROOT/$v python bench.py --no-cpu --no-configs --steps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'value', round(d['value']/1e9,3), 'e2e', round(d['e2e']['value']/1e9,3))"
  done
done

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --steps 1 --warmup 3 --no-cpu --frames 200000 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rx_frames -s 4 -c 1 -o gpurun_out/prof_rx python bench.py --steps 1 --warmup 3 --no-cpu --frames 200000 > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"; tail -2 gpurun_out/ncu2.log

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_stream_rx2 -s 4 -c 1 -o gpurun_out/prof_rx_full python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
python bench.py --extras > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"

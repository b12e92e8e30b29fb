cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu --frames 200000 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rx_frames -s 4 -c 2 -o gpurun_out/prof_rx python bench.py --steps 1 --warmup 3 --no-cpu --frames 200000 > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"; tail -3 gpurun_out/ncu2.log

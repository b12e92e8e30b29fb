cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu --frames 200000 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_stream_rx2 -s 4 -c 1 -o gpurun_out/prof_rx python bench.py --steps 1 --warmup 3 --no-cpu --frames 200000 > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"; tail -3 gpurun_out/ncu2.log

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/stream_probe.py 1000000 fast > gpurun_out/plain3.log 2>&1 && cat gpurun_out/plain3.log &&
ncu --set full --clock-control none --import-source on -k regex:'k_stream_rx2' -s 1 -c 1 -o gpurun_out/prof_stream python tools/stream_probe.py 1000000 fast > gpurun_out/ncu3.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu3.log

# round 2, call I (final library): guard probe, soak, configs[1] at full size against the compiled reference, bench --extras
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/guard_probe.py > gpurun_out/r2i_guard_probe.txt 2>&1; echo "guard rc=$?"; tail -8 gpurun_out/r2i_guard_probe.txt
timeout 900 python tools/checked_soak.py 4 6 > gpurun_out/r2i_soak.txt 2>&1; echo "soak rc=$?"; grep -E "decisions|all-SNR" gpurun_out/r2i_soak.txt | tail -8
timeout 600 python tools/full_parity.py --out gpurun_out/r2i_full_parity.json > gpurun_out/r2i_full_parity.txt 2>&1; echo "full parity rc=$?"; tail -5 gpurun_out/r2i_full_parity.txt
timeout 900 python bench.py --extras --no-cpu > gpurun_out/r2i_bench_extras.json 2> gpurun_out/r2i_bench_extras.err; echo "extras rc=$?"
timeout 300 python tools/replay_cost.py > gpurun_out/r2i_replay_cost.txt 2>&1; cat gpurun_out/r2i_replay_cost.txt

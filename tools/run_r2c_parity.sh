# configs[1] at its stated size against the compiled reference, on the shipped library
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/full_parity.py --out gpurun_out/r2c_full_parity.json > gpurun_out/r2c_full_parity.txt 2>&1; echo "full parity rc=$?"; tail -2 gpurun_out/r2c_full_parity.txt | cut -c1-260

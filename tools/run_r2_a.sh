# round 2, call A: parity of the packed-fp32 / speculated-channel kernels + A/B against the round-1 library
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt; nproc >> gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.txt
tail -5 gpurun_out/r2a_pytest.txt
timeout 600 bash tools/ab.sh build/ab/lib_r1.so build/ab/lib_p1.so > gpurun_out/r2a_ab.txt 2>&1
cat gpurun_out/r2a_ab.txt
timeout 600 python tools/full_parity.py --out gpurun_out/r2a_full_parity.json > gpurun_out/r2a_full_parity.txt 2>&1; echo "full_parity exit $?" >> gpurun_out/r2a_full_parity.txt
tail -4 gpurun_out/r2a_full_parity.txt

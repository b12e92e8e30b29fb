# the C driver alone on 8 GPUs (one process, one context per GPU, NCCL communicator warmed up before the clock starts)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
FRAMES_PER_GPU=16000000 bash tools/run_8gpu_driver.sh "1 8" 2>&1 | tee gpurun_out/r2b_c_driver_8gpu.txt
cd ieee-802.11-ofdm-qpsk-simulator_b200
mkdir -p /tmp/dN
for R in "" "--round-reduce-nccl"; do
  ./ofdm_sweep --quiet --outdir /tmp/dN --gpus 8 $R --target-errors 100 --max-bits 10000000000 --round-frames 4194304 --snr-start 0 --snr-count 21 --mode fast --seed 3 2>&1 | grep -v "NCCL version"
done | tee -a $GRAFT_REPO_ROOT/gpurun_out/r2b_c_driver_8gpu.txt

# the C driver's multi-GPU path (one process, one context per GPU, ncclAllReduce of the counters): configs[3] and configs[4]
# usage: run_8gpu_driver.sh "1 2 4 8"   (GPU counts to run, weak scaling: 16 M frames per GPU)
cd $GRAFT_REPO_ROOT/ieee-802.11-ofdm-qpsk-simulator_b200
mkdir -p /tmp/d
for N in ${1:-1 2 4 8}; do
  for TAPS in 0 8; do
    F=$((${FRAMES_PER_GPU:-16000000} * N))
    ./ofdm_sweep --quiet --outdir /tmp/d --gpus $N --frames $F --snr-start 0 --snr-count 21 --mode fast $( [ $TAPS -gt 0 ] && echo --taps $TAPS ) 2>&1 | grep -v "NCCL version"
    echo "   gpus $N taps $TAPS rc $?  BER(0,10,12 dB): $(cut -f1,11,13 /tmp/d/Output_BER.txt)"
  done
done

cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench$N.json 2> gpurun_out/bench$N.err; echo "N=$N rc=$?"; cat gpurun_out/bench$N.json | cut -c1-700; tail -2 gpurun_out/bench$N.err
done
cd ieee-802.11-ofdm-qpsk-simulator_b200 && mkdir -p /tmp/d && (time ./ofdm_sweep --quiet --outdir /tmp/d --gpus 8 --frames 80000000 --snr-start 0 --snr-count 21 --mode fast) 2>&1 | tail -4 && cat /tmp/d/Output_BER.txt

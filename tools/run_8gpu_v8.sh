cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(nvidia-smi topo -m; nproc; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)") > gpurun_out/topo8.txt 2>&1
for AFF in 0 1; do
OFDM_BENCH_AFFINITY=$AFF timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$AFF bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/v8_bench8_aff$AFF.json 2> gpurun_out/v8_bench8_aff$AFF.err; echo "aff=$AFF rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/v8_bench8_aff$AFF.json'))
print('value %.3e e2e %.3e frac %.3f affinity %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['config'].get('host_affinity')))"
done
timeout 300 python -m pytest tests -m gpu -q -k "two_gpu or multi or nccl" 2>&1 | tail -3

mkdir -p gpurun_out
for r in 16 8; do
echo "== nccl reserve $r"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2957$((r%10)) bench.py --gpus 8 --steps 6 --warmup 3 --no-e2e --no-cpu --concat-transport nccl --reserve-sms $r 2> gpurun_out/n8_$r.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'])"
done

mkdir -p gpurun_out
for r in 0 4 8; do
echo "== reserve $r"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$r bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu --reserve-sms $r 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'])"
done
echo "== overlap mode"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu --concat-mode overlap 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'])"

mkdir -p gpurun_out
echo "== peer"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --steps 6 --warmup 3 --no-e2e --no-cpu --concat-transport peer 2> gpurun_out/n8_peer.err | tee gpurun_out/bench_r01_octet_n8_peer.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'])"
grep -v "^\*\|OMP_NUM" gpurun_out/n8_peer.err | tail -3

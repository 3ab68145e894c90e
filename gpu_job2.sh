mkdir -p gpurun_out
for t in peer; do
echo "== $t"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e --no-cpu --concat-transport $t 2> gpurun_out/n2_$t.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'])"
grep -v "^\*\|OMP_NUM" gpurun_out/n2_$t.err | tail -4
done

mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_replay or flattened or erc20" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/n2.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'])"
grep -v "^\*\|OMP_NUM" gpurun_out/n2.err | tail -3

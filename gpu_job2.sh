mkdir -p gpurun_out
rm -f gpurun_out/n2_modes.txt
for mode in async simple; do for r in 0 2; do
echo "== $mode reserve $r" >> gpurun_out/n2_modes.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$r bench.py --gpus 2 --steps 8 --warmup 3 --no-e2e --no-cpu --reserve-sms $r --concat-mode $mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'])" >> gpurun_out/n2_modes.txt
done; done
cat gpurun_out/n2_modes.txt

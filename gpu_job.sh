mkdir -p gpurun_out
: > gpurun_out/n2_modes.jsonl
for mode in "--concat-mode simple" "--concat-mode overlap" "--concat-mode overlap --reserve-sms 8" "--concat-mode simple --reserve-sms 8" "--concat-mode overlap --reserve-sms 16"; do
  echo "{\"mode\": \"$mode\"}" >> gpurun_out/n2_modes.jsonl
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e $mode 2>> gpurun_out/n2_modes.err | grep '^{' >> gpurun_out/n2_modes.jsonl
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log

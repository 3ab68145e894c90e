# 2-GPU check of the final build: the multi-GPU parity tests and the bench under torchrun (push transport, self-verified)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multigpu.py tests/test_netting.py -q -m gpu > gpurun_out/pytest_multigpu.log 2>&1; tail -3 gpurun_out/pytest_multigpu.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err || tail -5 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2.json").read().strip().split("\n")[-1]); e = d.get("e2e") or {}; c = d.get("e2e_device_consumer") or {}
print("N=%d" % d["n_gpus"], round(d["value"] / 1e6, 1), "Mcyc/s", round(d["ms_per_step"], 3), "ms/step kernel_ms", round(d["roofline"]["kernel_ms"], 3),
      "| e2e", round(e.get("value", 0) / 1e6, 1), round(e.get("ms_per_step", 0), 1), e.get("host_ms_per_step"), "| consumer", round(c.get("value", 0) / 1e6, 1), "|", (d.get("multi_gpu") or {}).get("verified", "")[:60])
PY

# one default bench line (1 GPU) + its summary, after the codec parity tests
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_codec.py -q -m gpu -x 2>&1 | tail -1
timeout 400 python bench.py --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json")); e = d.get("e2e") or {}; c = d.get("e2e_device_consumer") or {}; r = d.get("e2e_raw_transport") or {}
print(round(d["value"] / 1e6, 1), "Mcyc/s", round(d["ms_per_step"], 3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"], 3), "frac", round(d["roofline"]["frac"], 4),
      "| e2e", round(e.get("value", 0) / 1e6, 1), round(e.get("ms_per_step", 0), 1), "ms subs", e.get("sub_batches"), "host", e.get("host_ms_per_step"),
      "| raw", round(r.get("value", 0) / 1e6, 1), "| consumer", round(c.get("value", 0) / 1e6, 1), round(c.get("ms_per_step", 0), 1))
PY

mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu --steps 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json")); e = d.get("e2e") or {}; c = d.get("e2e_device_consumer") or {}; r = d.get("e2e_raw_transport") or {}
print(round(d["value"] / 1e6, 1), "Mcyc/s | e2e", round(e.get("value", 0) / 1e6, 1), round(e.get("ms_per_step", 0), 1), "ms | raw", round(r.get("value", 0) / 1e6, 1), "| consumer", round(c.get("value", 0) / 1e6, 1), round(c.get("ms_per_step", 0), 1))
PY

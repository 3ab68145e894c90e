# fourth GPU job of round 2: single-pass transport encoder (staging + compaction): parity, kernel times, the bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_codec.py tests/test_host_replay.py tests/test_consumer.py -q -m gpu -x > gpurun_out/pytest_codec.log 2>&1; tail -5 gpurun_out/pytest_codec.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_aux.csv python tools/aux_kernels.py > gpurun_out/aux_launches.log 2>&1
grep -i "encode\|consume" gpurun_out/launches_r02_aux.csv | awk -F'","' '{print $5, $NF}' | cut -c1-160
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json")); e = d.get("e2e") or {}; c = d.get("e2e_device_consumer") or {}; r = d.get("e2e_raw_transport") or {}
print(round(d["value"] / 1e6, 1), "Mcyc/s", round(d["ms_per_step"], 3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"], 3), "frac", round(d["roofline"]["frac"], 4),
      "| e2e", round(e.get("value", 0) / 1e6, 1), round(e.get("ms_per_step", 0), 1), "ms d2h", e.get("d2h_bytes_per_step"), "ratio", e.get("ratio"), "host", e.get("host_ms_per_step"), "decode", e.get("host_decode"),
      "| raw", round(r.get("value", 0) / 1e6, 1), "| consumer", round(c.get("value", 0) / 1e6, 1), round(c.get("ms_per_step", 0), 1), c.get("host_ms_per_step"), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
PY

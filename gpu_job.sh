# The round's GPU validation job (run with: gpurun --timeout 1200 -- 'bash gpu_job.sh'):
# smoke, the whole -m gpu suite, the default bench line, the reference arm, the ncu launch list, and a --set full capture
# of the transport encoder / consumer / bytecode-hash kernels (summarised on the box: only text comes back).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
tail -4 gpurun_out/pytest_gpu_full.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json")); e = d.get("e2e") or {}; c = d.get("e2e_device_consumer") or {}; r = d.get("e2e_raw_transport") or {}
print(round(d["value"] / 1e6, 1), "Mcyc/s", round(d["ms_per_step"], 3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"], 3), "frac", round(d["roofline"]["frac"], 4),
      "| e2e", round(e.get("value", 0) / 1e6, 1), round(e.get("ms_per_step", 0), 1), "ms d2h", e.get("d2h_bytes_per_step"), "host", e.get("host_ms_per_step"),
      "| raw", round(r.get("value", 0) / 1e6, 1), "| consumer", round(c.get("value", 0) / 1e6, 1), round(c.get("ms_per_step", 0), 1), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
r = json.load(open("gpurun_out/bench_reference.json")); print("reference arm", round(r["value"] / 1e6, 1), "Mcyc/s", r["cpu_baseline"]["cores"], "cores")
PY
for wl in keccak storage; do
  timeout 150 python bench.py --no-cpu --no-e2e --steps 3 --workload $wl > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err || tail -2 gpurun_out/wl_$wl.err
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_bench.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_aux.csv python tools/aux_kernels.py > gpurun_out/aux_launches.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'zkb_(encode_kernel|encode_compact|consume|hash_bytecodes|logsort_gather)' -c 8 -o gpurun_out/ncu_r02_aux2 python tools/aux_kernels.py > gpurun_out/ncu_aux2.log 2>&1
python tools/ncu_multi_summary.py gpurun_out/ncu_r02_aux2.ncu-rep "ncu --set full --clock-control none: tools/aux_kernels.py (ERC-20 x8, 14 208 VMs): record gather of the radix sort, transport encoder (encoding pass + compaction), device-side consumer, bytecode hashing" > gpurun_out/ncu_r02_aux2.txt 2>&1
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out

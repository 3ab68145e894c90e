# third GPU job of round 2: codec v2 (CUDA encoder == scalar encoder), schedule regrouping, the full -m gpu suite, the bench
# line, and the ncu captures whose reports were lost in job 2 (summarised ON the box: only text + small reports come back)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_codec.py tests/test_host_replay.py -q -m gpu -x > gpurun_out/pytest_codec.log 2>&1; tail -15 gpurun_out/pytest_codec.log
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log; tail -4 gpurun_out/pytest_gpu_full.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
for wl in mixed mixed_shuffled; do
  timeout 150 python bench.py --no-cpu --no-e2e --steps 3 --workload $wl > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err || tail -2 gpurun_out/wl_$wl.err
done
ZKB_REGROUP=0 timeout 150 python bench.py --no-cpu --no-e2e --steps 3 --workload mixed > gpurun_out/wl_mixed_noregroup.json 2> gpurun_out/wl_mixed_noregroup.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/wl_*.json")) + ["gpurun_out/bench.json"]:
    try:
        d = json.load(open(f)); e = d.get("e2e") or {}; c = d.get("e2e_device_consumer") or {}; r = d.get("e2e_raw_transport") or {}
        print(f, round(d["value"] / 1e6, 1), "Mcyc/s", round(d["ms_per_step"], 3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"], 3), "frac", round(d["roofline"]["frac"], 4),
              "| e2e", round(e.get("value", 0) / 1e6, 1), round(e.get("ms_per_step", 0), 1), "ms d2h", e.get("d2h_bytes_per_step"), "ratio", e.get("ratio"), "host", e.get("host_ms_per_step"), "decode", e.get("host_decode"),
              "| raw", round(r.get("value", 0) / 1e6, 1), "| consumer", round(c.get("value", 0) / 1e6, 1), round(c.get("ms_per_step", 0), 1), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
# launch list (one pass per kernel) of the auxiliary kernels: per-kernel device time
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_aux.csv python tools/aux_kernels.py > gpurun_out/aux_launches.log 2>&1
# --set full captures, summarised here; the big reports do not travel
timeout 240 ncu --set full --clock-control none --import-source on -k regex:zkb_run_kernel -s 2 -c 2 -o gpurun_out/ncu_r02_keccak python bench.py --workload keccak --vms 14208 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_keccak.log 2>&1
python tools/ncu_multi_summary.py gpurun_out/ncu_r02_keccak.ncu-rep "ncu --set full --clock-control none: bench.py --workload keccak --vms 14208 (BASELINE config 3 shape: 8 keccak256 calls over a 4 KiB preimage per VM), the FAST and the FULL interpreter launch of one step" > gpurun_out/ncu_r02_keccak.txt 2>&1
timeout 240 ncu --set full --clock-control none -k regex:zkb_run_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_alu_loop python bench.py --workload alu_loop --vms 56832 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_alu_loop.log 2>&1
python tools/ncu_multi_summary.py gpurun_out/ncu_r02_alu_loop.ncu-rep "ncu --set full --clock-control none: bench.py --workload alu_loop --vms 56832 (register-only ADD / SUB / MUL loop, BASELINE config 1 shape at 4 full waves): integer-pipe utilisation of the U256 ALU inside the interpreter" > gpurun_out/ncu_r02_alu_loop.txt 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'zkb_(pack|flatten|encode_kernel|consume|logsort|hash|restore)' -c 30 -o gpurun_out/ncu_r02_aux python tools/aux_kernels.py > gpurun_out/ncu_aux.log 2>&1
python tools/ncu_multi_summary.py gpurun_out/ncu_r02_aux.ncu-rep "ncu --set full --clock-control none: tools/aux_kernels.py (ERC-20 x8, 14 208 VMs): pack, flatten, per-slot netting (radix sort), transport encoder, device-side consumer, bytecode hashing" > gpurun_out/ncu_r02_aux.txt 2>&1
rm -f gpurun_out/ncu_r02_aux.ncu-rep
for f in gpurun_out/*.ncu-rep; do [ $(stat -c %s $f) -gt 20000000 ] && rm -f $f; done
du -sh gpurun_out; ls -la gpurun_out | tail -25

# second GPU job of round 2: the other workloads' device legs, the integer micro-benchmark (row j3), ncu captures of the
# keccak / ALU-loop interpreter launches and of every auxiliary kernel, the pipelined end-to-end legs
mkdir -p gpurun_out
python tools/alu_microbench.py > gpurun_out/alu_microbench.json 2> gpurun_out/alu_microbench.err; cat gpurun_out/alu_microbench.json; tail -2 gpurun_out/alu_microbench.err
timeout 300 python bench.py --no-cpu > gpurun_out/bench_pipelined.json 2> gpurun_out/bench_pipelined.err; tail -2 gpurun_out/bench_pipelined.err
for wl in keccak storage alu_loop div_loop mixed mixed_shuffled; do
  timeout 150 python bench.py --no-cpu --no-e2e --steps 3 --workload $wl > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err || tail -2 gpurun_out/wl_$wl.err
done
timeout 120 python tools/aux_kernels.py 2> gpurun_out/aux_kernels_times.txt; cat gpurun_out/aux_kernels_times.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/wl_*.json")) + ["gpurun_out/bench_pipelined.json"]:
    try:
        d = json.load(open(f)); e = d.get("e2e") or {}; c = d.get("e2e_device_consumer") or {}
        print(f, round(d["value"] / 1e6, 1), "Mcyc/s", round(d["ms_per_step"], 3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"], 3), "frac", round(d["roofline"]["frac"], 4),
              "e2e", round(e.get("value", 0) / 1e6, 1), round(e.get("ms_per_step", 0), 1), "consumer", round(c.get("value", 0) / 1e6, 1), round(c.get("ms_per_step", 0), 1))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
timeout 240 ncu --set full --clock-control none --import-source on -k regex:zkb_run_kernel -s 2 -c 2 -o gpurun_out/ncu_r02_keccak python bench.py --workload keccak --vms 14208 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_keccak.log 2>&1
timeout 240 ncu --set full --clock-control none --import-source on -k regex:zkb_run_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_alu_loop python bench.py --workload alu_loop --vms 56832 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_alu_loop.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'zkb_(pack|flatten|encode_kernel|consume|logsort|hash|restore)' -c 40 -o gpurun_out/ncu_r02_aux python tools/aux_kernels.py > gpurun_out/ncu_aux.log 2>&1
ls -la gpurun_out | tail -20

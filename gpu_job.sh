mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
tail -4 gpurun_out/pytest_gpu_full.log
timeout 900 python bench.py > gpurun_out/bench_r01_octet.json 2> gpurun_out/bench_r01_octet.err; tail -2 gpurun_out/bench_r01_octet.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cat gpurun_out/bench_r01_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_octet.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.json 2>gpurun_out/ncu1.err
for w in storage keccak mixed alu_loop; do timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload $w 2>/dev/null > gpurun_out/bench_r01_octet_$w.json; done
python -c "
import json
d=json.load(open('gpurun_out/bench_r01_octet.json')); print(d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['frac'], d['cpu_baseline']['value']/1e6)
for w in ['storage','keccak','mixed','alu_loop']:
    d=json.load(open(f'gpurun_out/bench_r01_octet_{w}.json')); print(w, d['value']/1e6, d['roofline']['kernel_ms'], d['roofline']['frac'])
"

mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
tail -3 gpurun_out/pytest_gpu_full.log
timeout 900 python bench.py > gpurun_out/bench_r01_octet.json 2> gpurun_out/bench_r01_octet.err; tail -2 gpurun_out/bench_r01_octet.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r01_octet.json')); print(d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['frac'], d['cpu_baseline']['value']/1e6, d['gpu_launches'])"

# The round's GPU validation job (run with: gpurun --timeout 1500 -- 'bash gpu_job.sh'):
# smoke, the whole -m gpu suite, the default bench line, the reference arm.
mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | head -40; cat /sys/devices/system/node/node*/cpulist; nproc; free -g) > gpurun_out/host_info.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
tail -3 gpurun_out/pytest_gpu_full.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json

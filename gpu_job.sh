mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
tail -30 gpurun_out/pytest_parity.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_oct.json 2> gpurun_out/bench_oct.err; tail -3 gpurun_out/bench_oct.err; cat gpurun_out/bench_oct.json

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
timeout 900 python tools/fuzz_gpu.py 16 300 > gpurun_out/fuzz.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r01_d.json 2> gpurun_out/bench_r01_d.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err

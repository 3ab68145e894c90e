mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
timeout 900 python tools/fuzz_gpu.py 40 200 > gpurun_out/fuzz.log 2>&1

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
timeout 900 python bench.py > gpurun_out/bench_r01_g.json 2> gpurun_out/bench_r01_g.err

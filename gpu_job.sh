mkdir -p gpurun_out
timeout 900 python tools/fuzz_gpu.py 60 1000 > gpurun_out/fuzz_octet.log 2>&1; tail -5 gpurun_out/fuzz_octet.log

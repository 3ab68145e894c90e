mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
timeout 900 python bench.py > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err
timeout 600 python bench.py --workload keccak --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_keccak.json 2> gpurun_out/bench_keccak.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.json 2>gpurun_out/ncu1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zkb_run_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01_final python bench.py --vms 16384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2.json 2>gpurun_out/ncu2.err

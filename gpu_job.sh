mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zkb_run_kernel -s 3 -c 1 -f -o gpurun_out/prof_oct_keccak python bench.py --workload keccak --vms 28416 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu4.json 2>gpurun_out/ncu4.err

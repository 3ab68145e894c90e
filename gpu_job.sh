set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.json 2>gpurun_out/ncu1.err
ncu --set full --clock-control none --import-source on -k regex:zkb_run_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01_erc20 python bench.py --vms 16384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2.json 2>gpurun_out/ncu2.err
ls -la gpurun_out

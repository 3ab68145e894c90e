mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r01_octet.json 2> gpurun_out/bench_r01_octet.err; tail -2 gpurun_out/bench_r01_octet.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_octet.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.json 2>gpurun_out/ncu1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zkb_run_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01_octet python bench.py --vms 56832 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2.json 2>gpurun_out/ncu2.err
for w in storage keccak mixed alu_loop; do timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload $w 2>/dev/null > gpurun_out/bench_r01_octet_$w.json; done
ls -la gpurun_out | tail -12

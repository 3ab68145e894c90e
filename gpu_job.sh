# The round's GPU validation job (run with: gpurun --timeout 1200 -- 'bash gpu_job.sh'):
# smoke, the whole -m gpu suite, the default bench line, the ncu launch list and one --set full capture of the interpreter.
mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | head -40; cat /sys/devices/system/node/node*/cpulist; nproc; free -g) > gpurun_out/host_info.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log
tail -3 gpurun_out/pytest_gpu_full.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
# kernel variants on the headline workload (device leg only)
ZKB_BALANCE=1 timeout 120 python bench.py --no-cpu --no-e2e > gpurun_out/var_balance.json 2> gpurun_out/var_balance.err
ZKB_BALANCE=1 timeout 120 python bench.py --no-cpu --no-e2e --workload mixed --vms 131072 > gpurun_out/var_balance_mixed.json 2> gpurun_out/var_balance_mixed.err
timeout 120 python bench.py --no-cpu --no-e2e --workload mixed --vms 131072 > gpurun_out/var_main_mixed.json 2> gpurun_out/var_main_mixed.err
for t in p16 p64 p128; do
  ZKB_LIB_PATH=$PWD/build/variants/libzkb_$t.so timeout 120 python bench.py --no-cpu --no-e2e > gpurun_out/var_$t.json 2> gpurun_out/var_$t.err
done
ZKB_BALANCE=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/pytest_balance.log 2>&1; tail -1 gpurun_out/pytest_balance.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/var_*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["value"] / 1e6, 1), "Mcyc/s", round(d["ms_per_step"], 3), "ms; kernel_ms", round(d["roofline"]["kernel_ms"], 3), "frac", round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(f, "FAILED", e)
PY
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_bench.log 2>&1
# one --set full capture of the FAST interpreter launch of a warm step (4 full waves of 148 x 96 VMs)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:zkb_run_kernel -s 2 -c 1 -o gpurun_out/ncu_r02_erc20_56k python bench.py --vms 56832 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_erc20.log 2>&1
ls -la gpurun_out | tail -30

mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "snapshot or workload_parity" > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
tail -12 gpurun_out/pytest_parity.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
for w in erc20 mixed storage; do
echo "== sparse $w"; timeout 300 $B --workload $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'])"
echo "== dense $w"; ZKB_RESTORE_DENSE=1 timeout 300 $B --workload $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'])"
done

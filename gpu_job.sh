mkdir -p gpurun_out
ZKB_LIB_PATH=build/variants/libzkb_nosync.so timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
tail -3 gpurun_out/pytest_parity.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
for v in base nosync nosyncpin base nosync nosyncpin; do for w in erc20; do
  echo "== $v $w"; ZKB_LIB_PATH=build/variants/libzkb_$v.so timeout 300 $B --workload $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['kernel_ms'])"
done; done
for v in base nosync; do for w in mixed alu_loop; do
  echo "== $v $w"; ZKB_LIB_PATH=build/variants/libzkb_$v.so timeout 300 $B --workload $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['kernel_ms'])"
done; done

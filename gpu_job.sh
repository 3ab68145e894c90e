mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_bytecode_hash.py -x -q -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
tail -5 gpurun_out/pytest_parity.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
for w in erc20 alu_loop; do timeout 300 $B --workload $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['kernel_ms'])"; done
cat > /tmp/small_run.py <<'PY'
import sys
sys.path.insert(0, '.')
from era_zk_evm_b200 import GpuVmBatch, workloads
for w, n in ((workloads.Erc20(n_transfers=1), 8), (workloads.WORKLOADS["mixed"](n_programs=2), 64), (workloads.WORKLOADS["keccak"](n_calls=1, preimage_bytes=200), 4)):
    ids = list(range(n))
    b = GpuVmBatch(w.config(n)); w.setup(b, ids); b.run(); print(w.name, b.totals()[0], (b.vm_status()[:,0]==1).all())
PY
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 60 python /tmp/small_run.py > gpurun_out/racecheck.log 2>&1; grep -c "Race reported" gpurun_out/racecheck.log; tail -3 gpurun_out/racecheck.log

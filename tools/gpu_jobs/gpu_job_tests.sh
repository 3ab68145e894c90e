# the two parity tests added last: regrouped vs ungrouped schedule, unknown code hash through the hash index
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "regrouped or unknown_code" > gpurun_out/pytest_new.log 2>&1; tail -30 gpurun_out/pytest_new.log

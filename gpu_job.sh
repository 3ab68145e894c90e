mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
: > gpurun_out/variants.jsonl
for v in old w20 w24 w16; do
    echo "{\"variant\": \"$v\", \"schedule\": 2}" >> gpurun_out/variants.jsonl
    ZKB_LIB_PATH=$PWD/build/variants/libzkb_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zkb_run_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01_v4 python bench.py --vms 16384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2.json 2>gpurun_out/ncu2.err
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_full.log

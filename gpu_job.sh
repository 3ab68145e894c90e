mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
: > gpurun_out/variants.jsonl
for v in w20 w16 w24 w20noec w16noec; do
    echo "{\"variant\": \"$v\", \"schedule\": 2}" >> gpurun_out/variants.jsonl
    ZKB_LIB_PATH=$PWD/build/variants/libzkb_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r01_f.json 2> gpurun_out/bench_r01_f.err

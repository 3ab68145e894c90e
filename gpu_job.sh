set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
: > gpurun_out/variants.jsonl
for v in w16c1k1 w16c1k2 w16c1k4 w16c1k8 w16c1k16 w12c2k1 w12c2k2 w12c2k4 w12c2k8 w12c2k16 w24c1k4 w8c3k4; do
    echo "{\"variant\": \"$v\", \"schedule\": 2}" >> gpurun_out/variants.jsonl
    ZKB_LIB_PATH=$PWD/build/variants/libzkb_$v.so timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
done
ls -la gpurun_out

set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
ZKB_SCHEDULE=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_free.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_free.log
: > gpurun_out/variants.jsonl
for v in w16c1 w20c1 w24c1 w8c2 w10c2 w12c2 w8c3; do
  for s in 2; do
    echo "{\"variant\": \"$v\", \"schedule\": $s}" >> gpurun_out/variants.jsonl
    ZKB_LIB_PATH=$PWD/build/variants/libzkb_$v.so ZKB_SCHEDULE=$s timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
  done
done
ZKB_SCHEDULE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zkb_run_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01_v2_lockstep python bench.py --vms 16384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu2.json 2>gpurun_out/ncu2.err
ls -la gpurun_out

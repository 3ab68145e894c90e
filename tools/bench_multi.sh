#!/bin/bash
# usage: tools/bench_multi.sh <ngpus> <tag> [bench.py args...]  -> gpurun_out/multi_<tag>.json
n=$1; tag=$2; shift; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" > gpurun_out/multi_$tag.json 2> gpurun_out/multi_$tag.err || tail -5 gpurun_out/multi_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/multi_$tag.json").read().strip().split("\n")[-1]); e=d.get("e2e") or {}
    print("$tag", "N=%d"%d["n_gpus"], round(d["value"]/1e6,1), "Mcyc/s", round(d["ms_per_step"],3), "ms/step kernel_ms", round(d["roofline"]["kernel_ms"],3), "e2e", round(e.get("value",0)/1e6,1), (d.get("multi_gpu") or {}).get("verified","")[:20])
except Exception as ex: print("$tag FAILED", ex)
PY

#!/bin/bash
# usage: tools/bench_variants.sh "<tag> <tag> ..." "<workload> ..."   (tags: build/variants/libzkb_<tag>.so; "main" = the in-tree library)
mkdir -p gpurun_out
for tag in $1; do
  if [ "$tag" = "main" ]; then unset ZKB_LIB_PATH; else export ZKB_LIB_PATH=$PWD/build/variants/libzkb_$tag.so; fi
  for w in $2; do
    timeout 300 python bench.py --no-cpu --no-e2e --workload $w > gpurun_out/var_${tag}_$w.json 2> gpurun_out/var_${tag}_$w.err || tail -3 gpurun_out/var_${tag}_$w.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_${tag}_$w.json")); print("$tag $w", round(d["value"]/1e6,1), "Mcyc/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],3))
except Exception as e: print("$tag $w FAILED", e)
PY
  done
done

B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
for v in base pinlane pinsaddr base pinlane pinsaddr; do for w in erc20; do
  echo "== $v $w"; ZKB_LIB_PATH=build/variants/libzkb_$v.so timeout 300 $B --workload $w 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['roofline']['kernel_ms'])"
done; done

mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r01_octet_n$N.json 2> gpurun_out/bench_n$N.err; tail -2 gpurun_out/bench_n$N.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_r01_octet_n$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value']/1e6, d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e'] and d['e2e']['value']/1e6)"

nvidia-smi topo -m 2>&1 | head -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/peer_copy_bw.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -20

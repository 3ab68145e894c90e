#!/usr/bin/env python
"""2-rank micro-test: bandwidth of copying a device buffer into rank 0's IPC-mapped buffer (torch copy_ vs cudaMemcpyPeerAsync).
Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_copy_bw.py"""
import ctypes, os, time
import torch, torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n = 900 << 20
payload = [None]
if rank == 0:
    sink = torch.empty(2 * n, dtype=torch.uint8, device=dev)
    payload = [reduce_tensor(sink)]
dist.broadcast_object_list(payload, src=0)
if rank != 0:
    fn, args = payload[0]
    sink = fn(*args)
src = torch.full((n,), rank + 1, dtype=torch.uint8, device=dev)
print(rank, "sink device", sink.device, "can access peer", torch.cuda.can_device_access_peer(local, 0) if local != 0 else "-", flush=True)
dist.barrier()
if rank == 1:
    for name in ("torch copy_", "cudaMemcpyPeerAsync", "cudaMemcpyAsync default"):
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            if name == "torch copy_":
                sink[n:2 * n].copy_(src, non_blocking=True)
            else:
                rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
                st = torch.cuda.current_stream().cuda_stream
                if name == "cudaMemcpyPeerAsync":
                    rc = rt.cudaMemcpyPeerAsync(ctypes.c_void_p(sink.data_ptr() + n), 0, ctypes.c_void_p(src.data_ptr()), local, ctypes.c_size_t(n), ctypes.c_void_p(st))
                else:
                    rc = rt.cudaMemcpyAsync(ctypes.c_void_p(sink.data_ptr() + n), ctypes.c_void_p(src.data_ptr()), ctypes.c_size_t(n), 4, ctypes.c_void_p(st))
                assert rc == 0, rc
            torch.cuda.synchronize(); torch.cuda.synchronize(0)
            dt = time.perf_counter() - t0
            print(f"{name}: {n / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms)", flush=True)
# the product's own sink: cudaMalloc + IPC handle opened with the reader's device current (peer-mapped for kernels)
lib0 = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "era_zk_evm_b200", "libzkb.so"))
hbuf = (ctypes.c_uint8 * 64)()
ksink = ctypes.c_void_p()
if rank == 0:
    assert lib0.zkb_peer_sink_create(local, ctypes.c_uint64(2 * n), ctypes.byref(ksink), hbuf) == 0
hpay = [bytes(hbuf)]
dist.broadcast_object_list(hpay, src=0)
if rank != 0:
    hb = (ctypes.c_uint8 * 64).from_buffer_copy(hpay[0])
    rc = lib0.zkb_peer_sink_open(local, hb, ctypes.byref(ksink))
    assert rc == 0, rc
if rank == 1:
    lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "era_zk_evm_b200", "libzkb.so"))
    lib.zkb_peer_push_async.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p]
    for n_ctas in (4, 8, 16, 32, 64, 148, 296):
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            rc = lib.zkb_peer_push_async(local, 0, src.data_ptr(), ksink.value + n, n, n_ctas, torch.cuda.current_stream().cuda_stream)
            assert rc == 0, rc
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(f"zkb_peer_push_kernel {n_ctas} CTAs: {n / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms)", flush=True)
dist.barrier()
if rank == 0:
    torch.cuda.synchronize()
    print("rank0 sees", int(sink[n]), int(sink[2 * n - 1]))
dist.destroy_process_group()

"""-m gpu, needs >= 2 GPUs (skipped on a single-GPU box): the N > 1 concat of the per-GPU query-log streams on real
devices -- NCCL send/recv (shard.gather_many) and the one-sided NVLink push (shard.PeerSink over zkb_peer_push_async) --
must both deliver, on rank 0, exactly the rank-ordered concatenation of what the oracle emits for the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from era_zk_evm_b200 import GpuVmBatch, records, shard, workloads
        size_pg = dist.new_group(backend="gloo")
        w = workloads.Erc20(n_transfers=2)
        lo, hi = shard.partition(n_total, world, rank)
        ids = np.arange(lo, hi)
        b = GpuVmBatch(w.config(len(ids), device=rank))
        w.setup(b, ids)
        b.run()
        kinds = [records.STREAM_LOG, records.STREAM_DECOMMIT, records.STREAM_REFUND]
        stream = torch.cuda.current_stream().cuda_stream
        locals_ = [shard.device_bytes_as_tensor(*b.pack_stream_device_async(k, stream), dev) for k in kinds]
        nccl = [pg.wait() for pg in shard.gather_many(locals_, dst=0, size_group=size_pg)]
        cap = int(sum(int(t.numel()) for t in locals_) * 1.25) * world + (1 << 20)
        sink = shard.PeerSink(cap, dev, dst=0, size_group=size_pg)
        for attempt in range(3):                       # both buffers of the double-buffered sink, then the first again
            peer = [pg.wait() for pg in sink.gather_many(locals_)]
            torch.cuda.synchronize()
            if rank == 0:
                for (a, ao), (p, po) in zip(nccl, peer):
                    assert ao.tolist() == po.tolist()
                    assert torch.equal(a, p), f"peer push differs from the NCCL concat (attempt {attempt})"
        if rank == 0:
            np.savez(os.path.join(out_dir, "gathered.npz"), **{f"k{k}": t[0].cpu().numpy() for k, t in zip(kinds, nccl)})
        dist.barrier()
        sink.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_concat_nccl_and_peer_push_match_the_oracle(tmp_path, oracle_mod):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from era_zk_evm_b200 import records, workloads
    n_total, world = 150, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    w = workloads.Erc20(n_transfers=2)
    orc = oracle_mod.OracleBatch(w.config(n_total))
    w.setup(orc, np.arange(n_total))
    orc.run_threads(0, 0)
    for k in (records.STREAM_LOG, records.STREAM_DECOMMIT, records.STREAM_REFUND):
        want = np.concatenate([orc.read_stream(vm, k).view(np.uint8) for vm in range(n_total)])
        assert np.array_equal(got[f"k{k}"], want), records.STREAM_NAMES[k]


def _cabi_worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # only to hand the NCCL unique id around
    try:
        from era_zk_evm_b200 import GpuVmBatch, records, shard, workloads
        w = workloads.Erc20(n_transfers=2)
        lo, hi = shard.partition(n_total, world, rank)
        ids = np.arange(lo, hi)
        b = GpuVmBatch(w.config(len(ids), device=rank))
        w.setup(b, ids)
        b.run()
        comm = shard.Comm(dev)
        kinds = [records.STREAM_LOG, records.STREAM_DECOMMIT, records.STREAM_FRAME, records.STREAM_REFUND]
        saved = {}
        for step in range(2 * world):                          # the sink rotates over the ranks
            dst = step % world
            got = comm.gather_streams(b, kinds, dst)
            torch.cuda.synchronize()
            if rank == dst and step < world:
                for k, (t, offs) in got.items():
                    saved[f"g{dst}_k{k}"] = t.cpu().numpy()
                    saved[f"g{dst}_o{k}"] = offs
        for attempt in range(2):
            share, src = comm.exchange_logs(b)
            torch.cuda.synchronize()
            saved[f"x{attempt}"] = share.cpu().numpy()
            saved[f"xs{attempt}"] = src
        # the one-sided variant over NVLink peer memory: same shares (per source), same concat on the rotating sink
        small = [records.STREAM_DECOMMIT, records.STREAM_FRAME, records.STREAM_REFUND]
        for i in range(world + 1):
            dst = i % world
            tag = comm.push_step(b, small, dst)
            torch.cuda.synchronize()
            shares, concat = comm.push_result(b, tag)
            saved[f"p{i}"] = np.concatenate([t.cpu().numpy() for t in shares]) if shares else np.zeros(0, np.uint8)
            if rank == dst:
                for k in small:
                    saved[f"pc{i}_k{k}"] = np.concatenate([t.cpu().numpy() for t in concat[k]])
            dist.barrier()      # nobody starts the next step's pushes into buffers that are still being read
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **saved)
        dist.barrier()
        comm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_cabi_gather_and_hash_partitioned_exchange_match_the_oracle(tmp_path, oracle_mod):
    """zkb_gather_streams (rotating sink) and zkb_exchange_logs (balanced all-to-all by slot hash), NCCL driven from
    libzkb.so: every sink gets the rank-ordered concatenation of the oracle's streams; every rank gets exactly the
    oracle's LOG records whose slot hashes to it, in (rank, VM, position) order"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from era_zk_evm_b200 import records, shard, workloads
    n_total, world = 150, min(torch.cuda.device_count(), 4)
    mp.spawn(_cabi_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    w = workloads.Erc20(n_transfers=2)
    orc = oracle_mod.OracleBatch(w.config(n_total))
    w.setup(orc, np.arange(n_total))
    orc.run_threads(0, 0)
    got = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    for dst in range(world):
        for k in (records.STREAM_LOG, records.STREAM_DECOMMIT, records.STREAM_FRAME, records.STREAM_REFUND):
            want = np.concatenate([orc.read_stream(vm, k).view(np.uint8) for vm in range(n_total)])
            assert np.array_equal(got[dst][f"g{dst}_k{k}"], want), (dst, records.STREAM_NAMES[k])
            bounds = [shard.partition(n_total, world, r)[0] for r in range(world)] + [n_total]
            want_offs = [sum(orc.read_stream(vm, k).nbytes for vm in range(b)) for b in bounds]
            assert got[dst][f"g{dst}_o{k}"].tolist() == want_offs
    logs = np.concatenate([orc.read_stream(vm, records.STREAM_LOG) for vm in range(n_total)])
    dest = shard.log_destination(logs, world)
    assert len(set(dest.tolist())) == world                       # the partition really spreads
    for r in range(world):
        want = logs[dest == r]
        for attempt in range(2):
            assert got[r][f"x{attempt}"].tobytes() == want.tobytes(), f"rank {r} share (attempt {attempt})"
        for i in range(world + 1):
            assert got[r][f"p{i}"].tobytes() == want.tobytes(), f"rank {r} one-sided share (step {i})"
            if i % world == r:
                for k in (records.STREAM_DECOMMIT, records.STREAM_FRAME, records.STREAM_REFUND):
                    whole = np.concatenate([orc.read_stream(vm, k).view(np.uint8) for vm in range(n_total)])
                    assert np.array_equal(got[r][f"pc{i}_k{k}"], whole), (r, i, records.STREAM_NAMES[k])

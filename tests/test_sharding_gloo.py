"""CPU suite: the N>1 path (static VM-range partition + variable-length stream concatenation) with world_size 2
over gloo.  Each rank runs its shard of the same seeded workload on the oracle (host-resident stand-in for the
per-GPU batch); rank 0 checks the concatenation against a single-process run of the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_everything():
    from era_zk_evm_b200.shard import partition
    for n in (0, 1, 7, 8, 65536, 1000003):
        for world in (1, 2, 3, 8):
            spans = [partition(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from era_zk_evm_b200 import records, workloads
        from era_zk_evm_b200.shard import all_gather_varlen, gather_stream, partition
        w = workloads.StorageHeavy(n_iters=16)
        lo, hi = partition(n_total, world, rank)
        ids = np.arange(lo, hi)
        b = oracle.OracleBatch(w.config(len(ids)))
        w.setup(b, ids)
        b.run_threads(0, 1)
        result = {}
        for kind in (records.STREAM_LOG, records.STREAM_ROWS, records.STREAM_REFUND):
            out, offsets, counts = gather_stream(b, kind, dst=0)
            if rank == 0:
                result[kind] = (out.numpy().copy(), offsets)
            else:
                assert out is None
        # several streams with ONE size exchange (what bench.py does at N > 1)
        from era_zk_evm_b200.shard import gather_many
        locs = [torch.from_numpy(np.ascontiguousarray(np.concatenate(
            [b.read_stream(vm, k).view(np.uint8) for vm in range(b.n_vms)]))) for k in (records.STREAM_LOG, records.STREAM_REFUND)]
        many = [pg.wait() for pg in gather_many(locs, dst=0)]
        if rank == 0:
            assert np.array_equal(many[0][0].numpy(), result[records.STREAM_LOG][0])
            assert np.array_equal(many[1][0].numpy(), result[records.STREAM_REFUND][0])
        # the same with the sizes exchanged over a separate host-side group (bench.py's default at N > 1)
        size_pg = dist.new_group(backend="gloo")
        many2 = [pg.wait() for pg in gather_many(locs, dst=0, size_group=size_pg)]
        if rank == 0:
            assert np.array_equal(many2[0][0].numpy(), result[records.STREAM_LOG][0])
            assert np.array_equal(many2[1][0].numpy(), result[records.STREAM_REFUND][0])
            assert many2[0][1].tolist() == many[0][1].tolist()
        # ragged edge: an empty contribution from one rank, and the all-gather flavour
        local = torch.arange(5 * rank, dtype=torch.uint8)
        cat, offs = all_gather_varlen(local)
        assert cat.tolist() == [x % 256 for r in range(world) for x in range(5 * r)]
        assert offs.tolist() == [0] + list(np.cumsum([5 * r for r in range(world)]))
        if rank == 0:
            np.savez(os.path.join(out_dir, "gathered.npz"), **{f"k{k}": v[0] for k, v in result.items()},
                     **{f"o{k}": v[1] for k, v in result.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world_size_2_stream_concat_matches_single_process(tmp_path, oracle_mod):
    from era_zk_evm_b200 import records, workloads
    n_total, world = 37, 2            # odd on purpose: ragged shards
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(tmp_path, "gathered.npz"))
    w = workloads.StorageHeavy(n_iters=16)
    b = oracle_mod.OracleBatch(w.config(n_total))
    w.setup(b, np.arange(n_total))
    b.run_threads(0, 1)
    for kind in (records.STREAM_LOG, records.STREAM_ROWS, records.STREAM_REFUND):
        whole = np.concatenate([b.read_stream(vm, kind).view(np.uint8) for vm in range(n_total)])
        assert np.array_equal(got[f"k{kind}"], whole), records.STREAM_NAMES[kind]
        assert got[f"o{kind}"][-1] == whole.size

"""SURVEY §8 row f-2, second half: per-slot grouping of the storage history = InMemoryStorage::flatten_and_net_history().1
(/root/reference/src/testing/storage.rs:50-73).  CPU: the oracle's grouped output against a plain Python group-by of its
own flattened history (same slots, same queries per slot, history order kept).  -m gpu: the device radix sort
(zkb_net_storage_history / zkb_sort_log_queries, csrc/logsort.cuh) against the oracle, byte for byte."""
import numpy as np
import pytest

from era_zk_evm_b200 import workloads

CASES = [("storage", dict(n_iters=24), 40), ("erc20", dict(n_transfers=3), 130), ("mixed", dict(n_programs=12), 12 * 32)]


def _run(batch_cls, name, kwargs, n):
    w = workloads.WORKLOADS[name](**kwargs)
    b = batch_cls(w.config(n))
    w.setup(b, np.arange(n))
    b.run()
    b.flatten_logs()
    return b


def _slot(r):
    return (int(r["shard_id"]), r["address"].tobytes(), r["key"].tobytes())


@pytest.mark.parametrize("name,kwargs,n", CASES)
def test_oracle_grouping_is_a_group_by_of_its_history(name, kwargs, n, oracle_mod):
    b = _run(oracle_mod.OracleBatch, name, kwargs, n)
    recs, flags, offsets, n_slots = b.net_storage_history()
    assert int(flags.sum()) == n_slots
    saw_rollback = False
    for vm in range(n):
        hist = b.read_flat(vm, 0)
        want = {}
        for r in hist:
            want.setdefault(_slot(r), []).append(r.tobytes())
            saw_rollback |= bool(r["rollback"])
        lo, hi = int(offsets[vm]), int(offsets[vm + 1])
        assert hi - lo == len(hist)
        got, cur = {}, None
        for i in range(lo, hi):
            if flags[i]:
                cur = _slot(recs[i])
                assert cur not in got, "a slot was opened twice"
                got[cur] = []
            assert _slot(recs[i]) == cur
            got[cur].append(recs[i].tobytes())
        assert got == want
    assert saw_rollback or name == "erc20"


@pytest.mark.gpu
@pytest.mark.parametrize("name,kwargs,n", CASES + [("storage", dict(), 3000)])
def test_device_sort_matches_the_oracle(name, kwargs, n, oracle_mod):
    from era_zk_evm_b200 import GpuVmBatch
    g = _run(GpuVmBatch, name, kwargs, n)
    o = _run(oracle_mod.OracleBatch, name, kwargs, n)
    gr, gf, go, gs = g.net_storage_history()
    orr, of, oo, os_ = o.net_storage_history()
    assert go.tolist() == oo.tolist() and gs == os_
    assert gf.tobytes() == of.tobytes()
    assert gr.tobytes() == orr.tobytes()


@pytest.mark.gpu
def test_global_sort_of_a_log_stream(oracle_mod):
    """zkb_sort_log_queries without groups: one slot map over ALL VMs' LOG records (what the multi-GPU exchange feeds it):
    every slot contiguous, input order kept inside a slot"""
    import ctypes as C
    import torch
    from era_zk_evm_b200 import GpuVmBatch, load_library, records
    g = _run(GpuVmBatch, "erc20", dict(n_transfers=2), 300)
    buf, _ = g.fetch_stream_packed(records.STREAM_LOG)
    recs = buf.view(records.LOG_DTYPE)
    n = len(recs)
    d_in = torch.from_numpy(buf.copy()).cuda()
    d_out = torch.empty_like(d_in)
    d_flag = torch.empty(n, dtype=torch.uint8, device="cuda")
    lib = load_library()
    lib.zkb_sort_log_queries.argtypes = [C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]
    ng = C.c_uint64()
    assert lib.zkb_sort_log_queries(0, d_in.data_ptr(), n, None, d_out.data_ptr(), d_flag.data_ptr(), C.byref(ng), None) == 0
    out = d_out.cpu().numpy().view(records.LOG_DTYPE)
    flags = d_flag.cpu().numpy()
    want = {}
    for r in recs:
        want.setdefault((int(r["aux_byte"]) * 0, _slot(r)), []).append(r.tobytes())
    got, cur = {}, None
    for i in range(n):
        if flags[i]:
            cur = (0, _slot(out[i]))
            assert cur not in got
            got[cur] = []
        got[cur].append(out[i].tobytes())
    assert got == want and ng.value == len(want)

"""-m gpu: BASELINE.json's configurations at their full per-GPU sizes.  The oracle cannot run 10^5 VMs in test time,
so parity is checked (a) bit-exactly on a spread sample of VMs re-created on the oracle from their GLOBAL ids and
(b) on every VM through size-independent properties of the packed streams: cycle / timestamp arithmetic, per-cycle
record counts summing to the stream lengths, frame starts balancing frame ends, storage-write accounting."""
import numpy as np
import pytest

from era_zk_evm_b200 import records, workloads
from era_zk_evm_b200.isa import C

from parity_util import first_mismatch

pytestmark = pytest.mark.gpu


def run_full(w, n):
    from era_zk_evm_b200 import GpuVmBatch
    gpu = GpuVmBatch(w.config(n))
    w.setup(gpu, np.arange(n))
    gpu.run()
    return gpu


def check_sample_against_oracle(gpu, w, oracle_mod, sample):
    orc = oracle_mod.OracleBatch(w.config(len(sample)))
    w.setup(orc, np.asarray(sample))
    orc.run_threads(0, 0)
    gs, os_ = gpu.vm_status(), orc.vm_status()
    for j, vm in enumerate(sample):
        assert tuple(gs[vm]) == tuple(os_[j]), (vm, gs[vm], os_[j])
        for kind in range(records.N_STREAMS):
            m = first_mismatch(kind, gpu.read_stream(int(vm), kind), orc.read_stream(j, kind))
            assert m is None, f"vm {vm}: {m}"
        assert bytes(gpu.read_local_state(int(vm))) == bytes(orc.read_local_state(j)), vm
    orc.close()


def check_stream_invariants(gpu, expect_all_ended=True):
    n = gpu.n_vms
    st = gpu.vm_status()
    if expect_all_ended:
        assert (st[:, 0] == 1).all(), st[st[:, 0] != 1][:5]
    counts = {k: gpu.stream_counts(k).astype(np.int64) for k in range(records.N_STREAMS)}
    assert (counts[records.STREAM_ROWS] == st[:, 1]).all()              # one row per executed cycle
    buf, offsets = gpu.fetch_stream_packed(records.STREAM_ROWS)
    rows = buf.view(records.ROW_DTYPE)
    starts = (offsets[:-1] // records.RECORD_BYTES[0]).astype(np.int64)
    vm_of_row = np.repeat(np.arange(n), counts[records.STREAM_ROWS])
    local = np.arange(len(rows), dtype=np.int64) - starts[vm_of_row]
    assert (rows["cycle"] == local).all()                                # monotonic_cycle_counter (cycle.rs:411)
    assert (rows["timestamp"] == C.STARTING_TIMESTAMP + C.TIME_DELTA_PER_CYCLE * local).all()   # mod.rs:232-234
    # per-cycle record counts sum to the stream lengths, per VM
    for kind, per_row in ((records.STREAM_MEM, rows["n_mem"].astype(np.int64)),
                          (records.STREAM_LOG, rows["n_log"].astype(np.int64)),
                          (records.STREAM_DECOMMIT, (rows["n_dfr"] & 3).astype(np.int64)),
                          (records.STREAM_REFUND, ((rows["n_dfr"] >> 4) & 3).astype(np.int64))):
        assert (np.bincount(vm_of_row, weights=per_row, minlength=n).astype(np.int64) == counts[kind]).all(), kind
    frames_in_rows = np.bincount(vm_of_row, weights=((rows["n_dfr"] >> 2) & 3).astype(np.int64), minlength=n).astype(np.int64)
    assert (frames_in_rows + 1 == counts[records.STREAM_FRAME]).all()   # + the bootloader push
    # the last row of an ended VM leaves an empty callstack
    last = starts + counts[records.STREAM_ROWS] - 1
    ended = st[:, 0] == 1
    assert (rows["callstack_depth"][last[ended]] == 0).all()
    del rows, buf
    fbuf, _ = gpu.fetch_stream_packed(records.STREAM_FRAME)
    fr = fbuf.view(records.FRAME_DTYPE)
    vm_of_fr = np.repeat(np.arange(n), counts[records.STREAM_FRAME])
    balance = np.bincount(vm_of_fr, weights=np.where(fr["kind"] == 1, 1, -1), minlength=n)
    assert (balance[ended] == 0).all()                                   # every started context was finished
    return counts


def test_config2_erc20_65536_vms(oracle_mod):
    n = 65536
    w = workloads.Erc20(n_transfers=8)
    gpu = run_full(w, n)
    counts = check_stream_invariants(gpu)
    sample = sorted(set(np.linspace(0, n - 1, 48).astype(int).tolist() + [63, 127, 65535]))     # incl. reverting VMs (id % 64 == 63)
    check_sample_against_oracle(gpu, w, oracle_mod, sample)
    # storage accounting: every successful transfer performs exactly two SSTOREs; funded VMs (balance 2^132, amounts
    # < 2^128) complete all 8, broke VMs (balance 0, id % 64 == 63) revert every transfer before any write
    lbuf, loffs = gpu.fetch_stream_packed(records.STREAM_LOG)
    lg = lbuf.view(records.LOG_DTYPE)
    vm_of = np.repeat(np.arange(n), counts[records.STREAM_LOG])
    writes = np.bincount(vm_of, weights=((lg["aux_byte"] == C.STORAGE_AUX_BYTE) & (lg["rw_flag"] == 1)), minlength=n)
    broke = (np.arange(n) % 64) == 63
    assert (writes[~broke] == 16).all() and (writes[broke] == 0).all()
    assert (counts[records.STREAM_REFUND] == writes).all()               # one refund record per SSTORE (log.rs:99-102)
    gpu.close()


def test_config3_keccak_4k_65536_vms(oracle_mod):
    n = 65536
    w = workloads.KeccakHeavy(n_calls=2, preimage_bytes=4096)
    gpu = run_full(w, n)
    check_stream_invariants(gpu)
    check_sample_against_oracle(gpu, w, oracle_mod, np.linspace(0, n - 1, 24).astype(int).tolist() + [1, 2])
    gpu.close()


def test_config4_storage_32768_vms_per_gpu(oracle_mod):
    n = 32768
    w = workloads.StorageHeavy()
    gpu = run_full(w, n)
    check_stream_invariants(gpu)
    check_sample_against_oracle(gpu, w, oracle_mod, np.linspace(0, n - 1, 40).astype(int).tolist())
    gpu.close()


def test_config5_mixed_131072_vms_per_gpu(oracle_mod):
    n = 131072
    w = workloads.Mixed(n_programs=256)
    gpu = run_full(w, n)
    st = gpu.vm_status()
    assert (st[:, 0] == 1).all(), np.unique(st[:, 0], return_counts=True)
    check_stream_invariants(gpu)
    rng = np.random.RandomState(5)
    check_sample_against_oracle(gpu, w, oracle_mod, sorted(rng.choice(n, 96, replace=False).tolist()))
    gpu.close()

"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle, bit-exact, on every workload."""
import numpy as np
import pytest

from era_zk_evm_b200 import records, workloads

from parity_util import compare_batches

pytestmark = pytest.mark.gpu


def _pair(w, vm_ids, oracle_mod):
    from era_zk_evm_b200 import GpuVmBatch
    cfg = w.config(len(vm_ids))
    gpu = GpuVmBatch(cfg)
    orc = oracle_mod.OracleBatch(cfg)
    w.setup(gpu, vm_ids)
    w.setup(orc, vm_ids)
    return gpu, orc


@pytest.mark.parametrize("name,kwargs,n", [
    ("alu_loop", dict(cycles=1000), 8),
    ("alu_loop", dict(cycles=100), 300),
    ("div_loop", dict(cycles=300), 200),
    ("mixed_shuffled", dict(n_programs=24), 24 * 8 + 3),
    ("storage", dict(), 70),
    ("keccak", dict(n_calls=3), 40),
    ("keccak", dict(n_calls=2, preimage_bytes=200), 16),
    ("erc20", dict(n_transfers=4), 200),
    ("mixed", dict(n_programs=24), 24 * 32),
    ("mixed", dict(n_programs=16, seed=0xF00D), 16 * 32 + 5),
    ("mixed", dict(n_programs=16, seed=0xBEEF, target_cycles=900), 16 * 32),
])
def test_workload_parity(name, kwargs, n, oracle_mod):
    w = workloads.WORKLOADS[name](**kwargs)
    vm_ids = list(range(n))
    gpu, orc = _pair(w, vm_ids, oracle_mod)
    gpu.run()
    orc.run_threads(0, 0)
    st = gpu.vm_status()
    assert (st[:, 0] == 1).all(), f"not all VMs ended: {st[st[:, 0] != 1][:5]}"
    problems = compare_batches(gpu, orc)
    assert not problems, "\n".join(problems)
    gc, gb = gpu.totals()
    oc, ob = orc.totals()
    assert (gc, gb) == (oc, ob)


@pytest.mark.parametrize("name,kwargs,n", [
    ("storage", dict(), 70),
    ("erc20", dict(n_transfers=3), 130),
    ("mixed", dict(n_programs=24), 24 * 32),
])
def test_refund_aware_storage_oracle_parity(name, kwargs, n, oracle_mod):
    """with the refund-aware oracle on (ZkbConfig.reserved[1]) every SSTORE probes the slot's
    cold/warm marker first; streams (incl. the RefundRec type / value and the ergs they change) must match the oracle"""
    from era_zk_evm_b200 import GpuVmBatch
    w = workloads.WORKLOADS[name](**kwargs)
    vm_ids = list(range(n))
    cfg = w.config(n)
    cfg.reserved[1] = 40
    cfg.storage_slots = max(cfg.storage_slots, 128)      # reads of absent keys claim a slot for their marker
    gpu, orc = GpuVmBatch(cfg), oracle_mod.OracleBatch(cfg)
    w.setup(gpu, vm_ids)
    w.setup(orc, vm_ids)
    gpu.run()
    orc.run_threads(0, 0)
    problems = compare_batches(gpu, orc)
    assert not problems, "\n".join(problems)
    refunds = [r for vm in vm_ids[:40] for r in gpu.read_stream(vm, records.STREAM_REFUND)]
    assert any(int(r["refund_type"]) == 1 and int(r["refund_value"]) == 40 for r in refunds), "no repeated write in the sample"


@pytest.mark.parametrize("name,kwargs,n", [
    ("erc20", dict(n_transfers=3), 130),
    ("keccak", dict(n_calls=2, preimage_bytes=200), 16),
    ("mixed", dict(n_programs=16, seed=0xF00D), 16 * 32 + 5),
])
@pytest.mark.parametrize("cycles_before_snapshot", [0, 53])
def test_snapshot_restore_reruns_identically(name, kwargs, n, cycles_before_snapshot, oracle_mod):
    """zkb_snapshot / zkb_restore (VmLocalState + backends are plain Clone values in the reference, vm_state/mod.rs:53):
    the restore copies stack pages and heap slabs only up to their high-water marks, so a batch restored to a snapshot --
    taken before the run or in the middle of it -- must re-run to exactly the streams of the first run, i.e. the oracle's"""
    w = workloads.WORKLOADS[name](**kwargs)
    gpu, orc = _pair(w, list(range(n)), oracle_mod)
    if cycles_before_snapshot:
        gpu.run(max_cycles_per_vm=cycles_before_snapshot)
    gpu.snapshot()
    orc.run_threads(0, 0)
    for attempt in range(3):
        gpu.run()
        problems = compare_batches(gpu, orc)
        assert not problems, f"run {attempt}: " + "\n".join(problems)
        gpu.restore()
        if cycles_before_snapshot:
            st = gpu.vm_status()
            assert (st[:, 1] <= cycles_before_snapshot).all()       # back at the snapshot's cycle counts


@pytest.mark.parametrize("defer", [0, 1])
@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("kwargs,n", [(dict(n_calls=3), 40), (dict(n_calls=2, preimage_bytes=200), 16), (dict(n_calls=2, preimage_bytes=273), 101)])
def test_keccak_paths_give_identical_results(defer, schedule, kwargs, n, oracle_mod):
    """ZkbConfig.reserved[2] bit 0 selects the kernel with the deferred thread-per-state sponge for long keccak256 inputs;
    both kernels, under both schedules, must emit the oracle's bytes (digest patched into the heap word and the record)"""
    from era_zk_evm_b200 import GpuVmBatch
    w = workloads.KeccakHeavy(**kwargs)
    ids = list(range(n))
    cfg = w.config(n)
    cfg.reserved[2] = defer
    cfg.schedule = schedule
    gpu, orc = GpuVmBatch(cfg), oracle_mod.OracleBatch(cfg)
    w.setup(gpu, ids)
    w.setup(orc, ids)
    gpu.run()
    orc.run_threads(0, 0)
    problems = compare_batches(gpu, orc)
    assert not problems, "\n".join(problems)


def test_deferred_keccak_survives_resumed_runs(oracle_mod):
    """max_cycles slices that end right on / around the yielding cycle"""
    w = workloads.KeccakHeavy(n_calls=2)
    gpu, orc = _pair(w, list(range(9)), oracle_mod)
    for _ in range(400):
        gpu.run(max_cycles_per_vm=3)
        if gpu.execution_has_ended():
            break
    orc.run_threads(0, 0)
    problems = compare_batches(gpu, orc)
    assert not problems, "\n".join(problems)


def test_resumable_run_matches_single_run(oracle_mod):
    w = workloads.Erc20(n_transfers=2)
    gpu, orc = _pair(w, list(range(33)), oracle_mod)
    for _ in range(64):
        gpu.run(max_cycles_per_vm=37)
        if gpu.execution_has_ended():
            break
    orc.run_threads(0, 0)
    problems = compare_batches(gpu, orc)
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("schedule", [1, 2])
def test_schedules_give_identical_results(schedule, oracle_mod):
    """free-running and lockstep warp schedules are two launch shapes of the same interpreter"""
    w = workloads.Mixed(n_programs=8, seed=0xABCD)
    from era_zk_evm_b200 import GpuVmBatch
    ids = list(range(8 * 32 + 3))
    cfg = w.config(len(ids))
    cfg.schedule = schedule
    gpu, orc = GpuVmBatch(cfg), oracle_mod.OracleBatch(cfg)
    w.setup(gpu, ids)
    w.setup(orc, ids)
    gpu.run()
    orc.run_threads(0, 0)
    problems = compare_batches(gpu, orc)
    assert not problems, "\n".join(problems)


def test_alu_golden_vectors_on_gpu():
    """the CUDA U256 ALU against the committed Python-int golden vectors (no oracle involved)"""
    import test_oracle_golden as G
    from era_zk_evm_b200 import GpuVmBatch
    allc = G.load("u256_vectors.json")
    for op in sorted(G.OPS):
        G.check_alu_vectors(GpuVmBatch, op, [c for c in allc if c["op"] == op])


def test_keccak_golden_vectors_on_gpu():
    """keccak256 precompile on the GPU against the committed digests: the reference's 8 live test shapes
    (keccak256.rs:144-196) plus the 4 KiB / rate-boundary shapes, aligned and unaligned"""
    import test_oracle_golden as G
    import vm_harness as H
    from era_zk_evm_b200 import GpuVmBatch, isa
    from era_zk_evm_b200.asm import Code, Imm, Program, R, far_call_abi
    from era_zk_evm_b200.isa import C
    for v in G.load("hash_vectors.json")["keccak256"]:
        data, un = G.keccak_input(v), v["unalignment"]
        p = Program()
        p.const("abi", far_call_abi(0xFFFFFFFF, start=un, length=len(data)))
        p.add(Code("abi"), 0, 8)
        p.add(Imm(C.KECCAK256_PRECOMPILE_ADDRESS), 0, 7)
        p.far_call(R(8), 7, "fail")
        p.ld_ptr(R(1), 2)
        p.st(Imm(8000), 2)
        p.ret(isa.RET_OK, R(0))
        p.label("fail")
        p.ret(isa.RET_PANIC, R(0))
        heap = b"\xff" * un + data
        b = H.launch(GpuVmBatch, p, 3, contracts={C.KECCAK256_PRECOMPILE_ADDRESS: workloads.keccak_system_contract()},
                     heap=heap + b"\x00" * (-len(heap) % 32) if heap else None, heap_bound=8192,
                     cfg_over=dict(heap_bytes=8192 + 64))
        for vm in range(3):
            r = H.rows(b, vm)
            digest = [x for x in r if H.family_of(x) == "uma" and int(x["bits"]) & 0x40][0]
            assert H.val(digest["dst0"]).to_bytes(32, "big").hex() == v["digest"], v
        b.close()


def test_sha256_golden_vectors_on_gpu():
    import test_oracle_golden as G
    from era_zk_evm_b200 import GpuVmBatch
    G.check_sha256_precompile(GpuVmBatch)


def test_host_replay_on_gpu_streams():
    """include/zkb_host.hpp bound to libzkb.so: replay the GPU batch's streams through the C++ tracer mirror"""
    import os
    import test_host_replay as T
    from era_zk_evm_b200 import GpuVmBatch
    from era_zk_evm_b200.batch import LIB_PATH
    shim = T.build_shim("zkb_", LIB_PATH)
    w = workloads.Mixed(n_programs=8, seed=0x77)
    n = 8 * 32
    b = GpuVmBatch(w.config(n))
    spy = T.InitialStateSpy(b)
    w.setup(spy, np.arange(n))
    b.run()
    totals = T.replay_all(shim, b, spy, range(n))
    assert totals[0] == b.totals()[0]


def test_ecrecover_golden_vectors_on_gpu():
    """warp-cooperative secp256k1 recovery on the GPU against the reference's known answers, signatures produced by the
    `cryptography` package and the failure shapes (r/s out of range, x not on the curve)"""
    import test_oracle_golden as G
    from era_zk_evm_b200 import GpuVmBatch
    G.check_ecrecover_precompile(GpuVmBatch)


import semantic_cases  # noqa: E402


@pytest.mark.parametrize("case", semantic_cases.ALL, ids=lambda c: c.__name__)
def test_semantic_case_on_gpu(case):
    """the hand-derived quirk programs (expectations from the reference source) straight on the CUDA batch"""
    from era_zk_evm_b200 import GpuVmBatch
    case(GpuVmBatch)


@pytest.mark.parametrize("name,kwargs,n", [
    ("mixed", dict(n_programs=16, seed=0x51), 16 * 32),
    ("storage", dict(), 64),
    ("erc20", dict(n_transfers=3), 130),
])
def test_flattened_histories_match_the_oracle(name, kwargs, n, oracle_mod):
    """zkb_flatten_logs (GPU) vs the oracle's InMemoryStorage / InMemoryEventSink forward logs and flatten()"""
    w = workloads.WORKLOADS[name](**kwargs)
    gpu, orc = _pair(w, list(range(n)), oracle_mod)
    gpu.run()
    orc.run_threads(0, 0)
    gpu.flatten_logs()
    orc.flatten_logs()
    n_rollbacks = 0
    for kind in range(4):
        gc, gs = gpu.flat_counts(kind)
        oc, os_ = orc.flat_counts(kind)
        assert (gs == os_).all() and (gc == oc).all(), (kind, np.nonzero(gc != oc)[0][:5])
        for vm in range(n):
            a, b = gpu.read_flat(vm, kind), orc.read_flat(vm, kind)
            assert a.tobytes() == b.tobytes(), (kind, vm)
            n_rollbacks += int(a["rollback"].sum())
    if name != "erc20":
        assert n_rollbacks > 0


def test_regrouped_schedule_emits_the_same_bytes(oracle_mod, monkeypatch):
    """zkb_run schedules VMs in bootloader-code order (DESIGN.md §4 "Regrouping"); state and streams stay indexed by VM, so
    the grouped and the ungrouped (ZKB_REGROUP=0) schedule must produce identical streams -- and both equal the oracle's"""
    from era_zk_evm_b200 import GpuVmBatch
    w = workloads.WORKLOADS["mixed_shuffled"](n_programs=12)
    n = 12 * 9 + 5
    vm_ids = list(range(n))
    orc = oracle_mod.OracleBatch(w.config(n))
    w.setup(orc, vm_ids)
    orc.run_threads(0, 0)
    blobs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("ZKB_REGROUP", flag)            # read by zkb_create
        gpu = GpuVmBatch(w.config(n))
        w.setup(gpu, vm_ids)
        gpu.run()
        problems = compare_batches(gpu, orc)
        assert not problems, f"ZKB_REGROUP={flag}\n" + "\n".join(problems)
        blobs.append(gpu.fetch_encoded().tobytes())
        gpu.close()
    assert blobs[0] == blobs[1]


def test_unknown_code_hash_stops_the_vm_like_the_reference(oracle_mod):
    """the reference's only Err (decommitter.rs:50-56 -> far_call.rs:448): a far call to a well-formed code hash that was
    never loaded.  On the device the lookup goes through the hash index over the loaded bytecodes (a miss ends at an empty
    slot); status, cycle count and every record emitted before the stop must equal the oracle's"""
    from era_zk_evm_b200 import isa
    from era_zk_evm_b200._binding import storage_entries
    w = workloads.Erc20(n_transfers=2)
    n = 41
    gpu, orc = _pair(w, list(range(n)), oracle_mod)
    fake = int.from_bytes(bytes([1, 0, 0, 1]) + bytes(range(28)), "big")
    for b in (gpu, orc):
        b.populate_storage(storage_entries([(0, isa.C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, workloads.TOKEN_ADDRESS, fake)]), vm_lo=3, vm_hi=29)
    gpu.run()
    orc.run_threads(0, 0)
    gs, os_ = gpu.vm_status(), orc.vm_status()
    assert (gs == os_).all()
    assert (gs[3:29, 0] == 2).all() and (gs[:3, 0] == 1).all() and (gs[29:, 0] == 1).all()      # ZKB_VM_UNKNOWN_CODE_HASH
    problems = compare_batches(gpu, orc)
    assert not problems, "\n".join(problems)


"""The transport codec (include/zkb_codec.h).  CPU: decode(encode(streams)) == streams byte for byte on every workload,
truncated / corrupted blobs are rejected, the compression ratio on the ERC-20 workload is what DESIGN.md quotes.
-m gpu: the CUDA encoder's blob is bit-identical to the scalar encoder's over the oracle's streams, and decodes (on the
host, through libzkb.so's own zkb_decode_stream) to the oracle's records."""
import numpy as np
import pytest

from era_zk_evm_b200 import records, workloads
from era_zk_evm_b200._binding import EncodedWitness

CASES = [
    ("alu_loop", dict(cycles=100), 5),
    ("erc20", dict(n_transfers=3), 70),
    ("keccak", dict(n_calls=2, preimage_bytes=200), 6),
    ("storage", dict(n_iters=24), 9),
    ("mixed", dict(n_programs=12), 12 * 32 + 3),
]


def _oracle_run(oracle_mod, name, kwargs, n):
    w = workloads.WORKLOADS[name](**kwargs)
    b = oracle_mod.OracleBatch(w.config(n))
    w.setup(b, np.arange(n))
    b.run_threads(0, 0)
    return w, b


@pytest.mark.parametrize("name,kwargs,n", CASES)
def test_roundtrip_is_lossless(name, kwargs, n, oracle_mod):
    _, b = _oracle_run(oracle_mod, name, kwargs, n)
    blob = b.fetch_encoded()
    view = EncodedWitness(oracle_mod.lib(), "orc_", blob)
    st = b.vm_status()
    raw = 0
    for vm in range(n):
        c = view.counts(vm)
        assert tuple(c[6:8]) == tuple(st[vm])
        for kind in range(records.N_STREAMS):
            want = b.read_stream(vm, kind)
            got = view.read_stream(vm, kind)
            assert c[kind] == len(want)
            assert got.tobytes() == want.tobytes(), (vm, records.STREAM_NAMES[kind])
            raw += want.nbytes
    assert blob.size < raw                       # it does compress


@pytest.mark.parametrize("name,kwargs,n", CASES)
def test_fast_decoder_matches_the_reference_decoder(name, kwargs, n, oracle_mod):
    """libzkb.so decodes with block predictions + a walk over the mask bits (zkb_codec::decode_records); the oracle
    library keeps the word-by-word inverse of the encoder (decode_records_ref).  Same bytes on every stream of every VM."""
    from era_zk_evm_b200 import load_library
    _, b = _oracle_run(oracle_mod, name, kwargs, n)
    blob = b.fetch_encoded()
    fast, ref = EncodedWitness(load_library(), "zkb_", blob), EncodedWitness(oracle_mod.lib(), "orc_", blob)
    for vm in range(n):
        for kind in range(records.N_STREAMS):
            a = fast.read_stream(vm, kind)
            assert a.tobytes() == ref.read_stream(vm, kind).tobytes() == b.read_stream(vm, kind).tobytes(), (vm, kind)
    # truncated slices are rejected by the fast decoder too
    from era_zk_evm_b200._binding import ZkbError
    with pytest.raises(ZkbError):
        EncodedWitness(load_library(), "zkb_", blob[: blob.size - 64]).read_stream(n - 1, records.STREAM_REFUND if name == "storage" else 0)


def test_bulk_decode_matches_per_vm_decode(oracle_mod):
    """zkb_decode_all (multi-threaded, the product's host decoder in libzkb.so -- no GPU needed) == per-VM streams"""
    from era_zk_evm_b200 import load_library
    import ctypes as C
    _, b = _oracle_run(oracle_mod, "erc20", dict(n_transfers=2), 50)
    blob = b.fetch_encoded()
    lib = load_library()
    lib.zkb_decode_all.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
    for kind in range(records.N_STREAMS):
        want = np.concatenate([b.read_stream(vm, kind).view(np.uint8) for vm in range(50)])
        for threads in (1, 3, 0):
            offsets = np.zeros(51, dtype=np.uint64)
            assert lib.zkb_decode_all(blob.ctypes.data, blob.size, kind, None, 0, offsets.ctypes.data, threads) == 0
            out = np.zeros(int(offsets[-1]), dtype=np.uint8)
            assert lib.zkb_decode_all(blob.ctypes.data, blob.size, kind, out.ctypes.data, out.size, offsets.ctypes.data, threads) == 0
            assert out.tobytes() == want.tobytes()
        assert lib.zkb_decode_all(blob.ctypes.data, blob.size, kind, out.ctypes.data, max(out.size, 1) - 1, offsets.ctypes.data, 1) != 0 or out.size == 0


def test_empty_and_unstarted_vms_encode(oracle_mod):
    """ragged input: VMs that never ran (no bootloader frame, zero records) next to finished ones"""
    w = workloads.AluLoop(cycles=100)
    b = oracle_mod.OracleBatch(w.config(4))
    blob0 = b.fetch_encoded()                    # nothing populated at all
    v0 = EncodedWitness(oracle_mod.lib(), "orc_", blob0)
    assert all(len(v0.read_stream(vm, k)) == 0 for vm in range(4) for k in range(6))
    w.setup(b, np.arange(4))
    b.run_threads(17, 1)                          # stopped in the middle of the loop
    v1 = EncodedWitness(oracle_mod.lib(), "orc_", b.fetch_encoded())
    for vm in range(4):
        assert v1.read_stream(vm, 0).tobytes() == b.read_stream(vm, 0).tobytes()
        assert len(v1.read_stream(vm, 0)) == 17


def test_malformed_blobs_are_rejected(oracle_mod):
    _, b = _oracle_run(oracle_mod, "erc20", dict(n_transfers=1), 3)
    blob = b.fetch_encoded().copy()
    lib = oracle_mod.lib()
    from era_zk_evm_b200._binding import ZkbError
    with pytest.raises(ZkbError):
        EncodedWitness(lib, "orc_", blob[: blob.size // 2]).read_stream(0, 0)          # truncated
    bad = blob.copy()
    bad[0] ^= 0xFF                                                                      # magic
    with pytest.raises(ZkbError):
        EncodedWitness(lib, "orc_", bad).counts(0)
    with pytest.raises(ZkbError):
        EncodedWitness(lib, "orc_", blob).counts(3)                                     # VM out of range
    # a presence bitmap that claims more residual words than the VM's slice holds
    hdr = blob[:128].view(np.uint64)
    payload0 = int(hdr[5])
    bad = blob.copy()
    bad[payload0:payload0 + 8] = 0xFF
    with pytest.raises(ZkbError):
        EncodedWitness(lib, "orc_", bad).read_stream(0, 0)


def test_corrupted_payloads_never_crash_the_decoder(oracle_mod):
    """the host decoder runs on bytes that crossed a link: flipped payload bytes must end in an error (or in different records
    when the flip hit a residual), never in a read or write outside the buffers.  Both decoders, every stream."""
    from era_zk_evm_b200 import load_library
    from era_zk_evm_b200._binding import ZkbError
    _, b = _oracle_run(oracle_mod, "erc20", dict(n_transfers=2), 12)
    blob = b.fetch_encoded()
    hdr = blob[:128].view(np.uint64)
    payload_lo, total = int(hdr[6]), int(hdr[2])          # payload_offset[0], total_bytes
    rng = np.random.default_rng(1234)
    rejected = 0
    for trial in range(120):
        bad = blob.copy()
        for at in rng.integers(payload_lo, total, size=int(rng.integers(1, 6))):
            bad[at] = rng.integers(0, 256)
        for lib, prefix in ((load_library(), "zkb_"), (oracle_mod.lib(), "orc_")):
            view = EncodedWitness(lib, prefix, bad)
            for vm in range(12):
                for kind in range(records.N_STREAMS):
                    try:
                        view.read_stream(vm, kind)
                    except ZkbError:
                        rejected += 1
    assert rejected > 0            # (most flips land in a presence mask and are caught by the length checks)


def test_decoders_under_address_sanitizer(oracle_mod, tmp_path):
    """tests/fuzz_codec.cpp built with -fsanitize=address,undefined: a few hundred corrupted blobs (payload bytes, every
    fourth trial also the count / offset tables) through both decoders of include/zkb_codec.h.  Any read or write outside a
    buffer aborts the harness.  (Found one: a 12-word DECOMMIT record carries a 32-bit mask word, and presence bits beyond
    word 11 used to be applied.)"""
    import os
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "fuzz_codec"
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", str(exe),
                        os.path.join(root, "tests", "fuzz_codec.cpp")], capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr or "").lower():
        pytest.skip("no sanitizer runtime in this image")
    assert r.returncode == 0, r.stderr
    for name, kwargs, n in (("erc20", dict(n_transfers=2), 12), ("mixed", dict(n_programs=6), 40)):
        _, b = _oracle_run(oracle_mod, name, kwargs, n)
        path = tmp_path / f"{name}.bin"
        b.fetch_encoded().tofile(str(path))
        out = subprocess.run([str(exe), str(path), "300"], capture_output=True, text=True)
        assert out.returncode == 0 and "decodes ok" in out.stdout, out.stderr[-2000:]


def test_wire_format_is_pinned(oracle_mod):
    """tests/golden/codec_v2.json: the blob of three small seeded workloads, byte for byte (sha256).  The CUDA encoder is
    checked against the scalar encoder on the GPU; this pins the scalar encoder -- i.e. the wire format -- itself."""
    import hashlib
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "codec_v2.json")) as f:
        golden = json.load(f)
    for name, case in golden["cases"].items():
        _, b = _oracle_run(oracle_mod, name, case["kwargs"], case["n_vms"])
        blob = b.fetch_encoded()
        assert blob.size == case["blob_bytes"] and sum(b.totals()[1]) == case["raw_bytes"], name
        assert hashlib.sha256(blob.tobytes()).hexdigest() == case["blob_sha256"], name


def test_erc20_ratio(oracle_mod):
    """the figure DESIGN.md / bench.py quote: the blob is ~17 % of the canonical bytes on the ERC-20 workload (format v2)"""
    _, b = _oracle_run(oracle_mod, "erc20", dict(n_transfers=8), 64)
    blob = b.fetch_encoded()
    raw = sum(b.totals()[1])
    assert 0.12 < blob.size / raw < 0.20, blob.size / raw


@pytest.mark.gpu
@pytest.mark.parametrize("name,kwargs,n", CASES + [("erc20", dict(n_transfers=8), 1000)])
def test_cuda_encoder_matches_the_scalar_encoder(name, kwargs, n, oracle_mod):
    from era_zk_evm_b200 import GpuVmBatch, load_library
    w, orc = _oracle_run(oracle_mod, name, kwargs, n)
    gpu = GpuVmBatch(w.config(n))
    w.setup(gpu, np.arange(n))
    gpu.run()
    blob_gpu, blob_orc = gpu.fetch_encoded(), orc.fetch_encoded()
    assert blob_gpu.size == blob_orc.size
    assert blob_gpu.tobytes() == blob_orc.tobytes()
    view = EncodedWitness(load_library(), "zkb_", blob_gpu)       # the product's own host decoder
    for vm in list(range(min(n, 40))) + [n - 1]:
        for kind in range(records.N_STREAMS):
            assert view.read_stream(vm, kind).tobytes() == orc.read_stream(vm, kind).tobytes()


@pytest.mark.gpu
def test_cuda_encoder_on_a_resumed_run(oracle_mod):
    """streams are cumulative across zkb_run calls: encoding after a partial run and after the final one"""
    from era_zk_evm_b200 import GpuVmBatch
    w = workloads.Erc20(n_transfers=2)
    n = 37
    gpu, orc = GpuVmBatch(w.config(n)), oracle_mod.OracleBatch(w.config(n))
    w.setup(gpu, np.arange(n))
    w.setup(orc, np.arange(n))
    gpu.run(max_cycles_per_vm=41)
    orc.run_threads(41, 1)
    assert gpu.fetch_encoded().tobytes() == orc.fetch_encoded().tobytes()
    gpu.run()
    orc.run_threads(0, 1)
    assert gpu.fetch_encoded().tobytes() == orc.fetch_encoded().tobytes()


@pytest.mark.gpu
def test_subset_blob_carries_only_the_requested_streams(oracle_mod):
    """zkb_fetch_encoded_kinds_async: what a host that consumes on the device still downloads (the query logs)"""
    import torch
    from era_zk_evm_b200 import GpuVmBatch, load_library
    w, orc = _oracle_run(oracle_mod, "erc20", dict(n_transfers=2), 60)
    gpu = GpuVmBatch(w.config(60))
    w.setup(gpu, np.arange(60))
    gpu.run()
    kinds = [records.STREAM_LOG, records.STREAM_DECOMMIT, records.STREAM_FRAME, records.STREAM_REFUND]
    full = gpu.fetch_encoded()
    pinned = torch.empty(full.size, dtype=torch.uint8, pin_memory=True)
    n = gpu.fetch_encoded_kinds_async(kinds, pinned.data_ptr(), pinned.numel())
    torch.cuda.synchronize()
    assert n < full.size // 3
    view = EncodedWitness(load_library(), "zkb_", pinned[:n].numpy())
    for vm in (0, 17, 59):
        for kind in range(records.N_STREAMS):
            got = view.read_stream(vm, kind)
            if kind in kinds:
                assert got.tobytes() == orc.read_stream(vm, kind).tobytes()
            else:
                assert len(got) == 0
        assert view.counts(vm)[0] == len(orc.read_stream(vm, 0))       # the true counts are still reported


def test_midcycle_stop_round_trips(oracle_mod):
    """a VM whose far call hits an unknown code hash stops in the MIDDLE of a cycle: records of that cycle without a row"""
    from era_zk_evm_b200 import isa
    from era_zk_evm_b200._binding import storage_entries
    w = workloads.Erc20(n_transfers=2)
    n = 9
    b = oracle_mod.OracleBatch(w.config(n))
    w.setup(b, np.arange(n))
    fake = int.from_bytes(bytes([1, 0, 0, 1]) + bytes(range(28)), "big")       # well-formed versioned hash, never loaded
    b.populate_storage(storage_entries([(0, isa.C.DEPLOYER_SYSTEM_CONTRACT_ADDRESS, workloads.TOKEN_ADDRESS, fake)]), vm_lo=0, vm_hi=5)
    b.run_threads(0, 1)
    st = b.vm_status()
    assert (st[:5, 0] == 2).all() and (st[5:, 0] == 1).all()                     # ZKB_VM_UNKNOWN_CODE_HASH / ended
    view = EncodedWitness(oracle_mod.lib(), "orc_", b.fetch_encoded())
    for vm in range(n):
        for kind in range(records.N_STREAMS):
            assert view.read_stream(vm, kind).tobytes() == b.read_stream(vm, kind).tobytes()


@pytest.mark.gpu
def test_capacity_stopped_batch_round_trips_on_the_gpu():
    """every VM stopped by a stream capacity (ZKB_VM_CAP_STREAM at row emission): whatever the batch reports as its streams,
    the CUDA encoder's blob must decode (host decoder) to exactly those canonical records"""
    from era_zk_evm_b200 import GpuVmBatch, load_library
    w = workloads.Erc20(n_transfers=2)
    n = 40
    cfg = w.config(n)
    cfg.cap_records[0] = 37                                                       # rows: every VM stops on ZKB_VM_CAP_STREAM
    gpu = GpuVmBatch(cfg)
    w.setup(gpu, np.arange(n))
    gpu.run()
    st = gpu.vm_status()
    assert (st[:, 0] >= 16).all(), st[:3]                                        # ZKB_VM_CAP_*
    view = EncodedWitness(load_library(), "zkb_", gpu.fetch_encoded())
    for vm in range(n):
        for kind in range(records.N_STREAMS):
            assert view.read_stream(vm, kind).tobytes() == gpu.read_stream(vm, kind).tobytes(), (vm, kind)

